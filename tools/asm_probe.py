"""Developer probe: assembly kernel (new fast records vs FEMB200_ASM_OLD=1) on the n=1448 P2 workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fem-libraries_b200")]
import numpy as np, torch
from femb200 import fem, mesh as fm

def timeit(fn, k=20, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(k): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / k

for kind, n in [(k, int(v)) for k, v in (a.split(":") for a in os.environ.get("CASES", "P2:1448,P1:2048").split(","))]:
    m = fm.jitter(fm.structured_triangles(n, order=1 if kind == "P1" else 2), 0.2, seed=1234)
    E = fm.young_per_cell(m.ncells)
    form = fem.ElasticityForm(m, E)
    A = fem.create_matrix(form)
    nb = 8 * A.nnz + 32 * m.ncells + 16 * m.nnodes
    res = {}
    for tag, env in (("old", {"FEMB200_ASM_OLD": "1"}), ("new", {})):
        for k in ("FEMB200_ASM_OLD", "FEMB200_ASM_NOSTREAM"): os.environ.pop(k, None)
        os.environ.update(env)
        t = timeit(lambda: fem.assemble_matrix(A, form))
        res[tag] = A.values.clone()
        print(f"{kind} n={n} {tag}: {t:.3f} ms  {m.ndofs / t / 1e6:.2f} GDOF/s  frac {nb / t / 1e6 / 6451.2:.3f}  plan {A.plan_bytes/1e6:.0f} MB", flush=True)
    print("   max |new - old| =", (res["new"] - res["old"]).abs().max().item())
