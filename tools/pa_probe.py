"""Developer probe: matrix-free (partial assembly) apply, config 3 (Q2 quads)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fem-libraries_b200")]
import numpy as np, torch
from femb200 import fem, mesh as fm

def timeit(fn, k=10, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(k): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / k

for kind, n in [(k, int(v)) for k, v in (a.split(":") for a in os.environ.get("CASES", "Q2:2048,Q2:4096,P2:1448").split(","))]:
    t0 = time.time()
    m = fm.structured_quads_q2(n) if kind == "Q2" else fm.structured_triangles(n, order=2)
    m = fm.jitter(m, 0.2, seed=1234)
    E = fm.young_per_cell(m.ncells)
    bc, g = fm.dirichlet_markers(m)
    form = fem.ElasticityForm(m, E)
    pa = fem.PAOperator(form, bcs=[fem.DirichletBC(bc, g)])
    v = torch.randn(m.ndofs, dtype=torch.float64, device="cuda")
    y = torch.empty_like(v)
    t = timeit(lambda: pa.mult(v, y))
    nv = m.xdofmap.shape[1]
    nbytes = m.nnodes * 32 + m.ncells * (8 * (2 * nv + 2) + 4 * m.nd + 4)
    print(f"{kind} n={n}: {m.ncells} cells {m.ndofs} dofs; pa_apply {t:.3f} ms = {m.ndofs/t/1e6:.1f} GDOF/s, "
          f"{nbytes/t/1e6:.0f} GB/s ({nbytes/t/1e6/6451.2:.3f} of roofline, {nbytes/m.ncells:.0f} B/cell); setup {time.time()-t0:.0f}s", flush=True)
    cg = fem.CGSolver(rel_tol=0.0, max_iter=10)
    cg.SetOperator(pa); cg.SetPreconditioner("jacobi")
    b = torch.ones_like(v)
    tc = timeit(lambda: cg.Mult(b, y, fixed_iters=10), 3, 1) / 11
    print(f"   PA-CG iteration {tc:.3f} ms = {m.ndofs/tc/1e6:.1f} GDOF/s", flush=True)
    del pa, form, v, y, b, cg
    torch.cuda.empty_cache()
