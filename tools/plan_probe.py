"""Developer probe: wall time of create_matrix (pattern + plan build) and of PAOperator set-up."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fem-libraries_b200")]
import numpy as np, torch
from femb200 import fem, mesh as fm

n = int(os.environ.get("N", "1448"))
m = fm.jitter(fm.structured_triangles(n, order=2), 0.2, seed=1234)
E = fm.young_per_cell(m.ncells)
form = fem.ElasticityForm(m, E)
tiny = fm.structured_triangles(4, order=2)
fem.create_matrix(fem.ElasticityForm(tiny, fm.young_per_cell(tiny.ncells)))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    A = fem.create_matrix(form)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"create_matrix #{rep}: {1e3*(t1-t0):.1f} ms, plan {A.plan_bytes/1e6:.0f} MB", flush=True)
    del A
    torch.cuda.synchronize(); print(f"   destroy: {1e3*(time.perf_counter()-t1):.1f} ms")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pa = fem.PAOperator(form)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"PAOperator #{rep}: {1e3*(t1-t0):.1f} ms", flush=True)
    del pa
