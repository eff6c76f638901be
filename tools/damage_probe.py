"""Developer probe: repeated tangent reassembly with damage (config 5 workload on one GPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fem-libraries_b200")]
import numpy as np, torch
from femb200 import fem, mesh as fm

def timeit(fn, k=5, w=2):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(k): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / k

n = int(os.environ.get("N", "1448"))
m = fm.jitter(fm.structured_triangles(n, order=2), 0.2, seed=1234)
E = fm.young_per_cell(m.ncells)
d = fm.damage_band(m)
u = 1e-3 * np.random.default_rng(0).standard_normal(m.ndofs)
frac = float((d[m.xdofmap].mean(axis=1) > 0).mean())
print(f"P2 n={n}: {m.ncells} cells, damaged cells {100*frac:.1f} %")
for variant, name in ((0, "closed form"), (1, "AD (nested duals)")):
    form = fem.ElasticityForm(m, E, 0.3, d=d, u=u, variant=variant)
    A = fem.create_matrix(form)
    t = timeit(lambda: fem.assemble_matrix(A, form))
    print(f"reassembly, damaged tangent {name}: {t:.3f} ms = {m.ndofs/t/1e6:.2f} GDOF/s", flush=True)
form = fem.ElasticityForm(m, E, 0.3)
A = fem.create_matrix(form)
t = timeit(lambda: fem.assemble_matrix(A, form))
print(f"linear (d = 0) fast path: {t:.3f} ms")
os.environ["FEMB200_FORCE_GENERIC"] = "1"
t = timeit(lambda: fem.assemble_matrix(A, form))
print(f"linear through the generic path: {t:.3f} ms")
