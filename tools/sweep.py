"""Developer sweep: assembly / SpMV kernel variants on the n=1448 P2 workload (GPU box)."""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fem-libraries_b200")]
import numpy as np
import torch
from femb200 import fem, mesh as fm

n = int(os.environ.get("N", "1448"))
kind = os.environ.get("KIND", "P2")
m = {"P1": lambda: fm.structured_triangles(n, order=1), "P2": lambda: fm.structured_triangles(n, order=2),
     "Q2": lambda: fm.structured_quads_q2(n)}[kind]()
m = fm.jitter(m, 0.2, seed=1234)
E = fm.young_per_cell(m.ncells)
form = fem.ElasticityForm(m, E)
A = fem.create_matrix(form)
print(kind, "n", n, "cells", m.ncells, "nnz", A.nnz, flush=True)


def timeit(fn, k=10, w=3):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(k):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / k


ref = None
for tpn, ch, R in itertools.product(os.environ.get("TPNS", "1,2").split(","), os.environ.get("CHS", "1,2,3").split(","), os.environ.get("RS", "64,96,128").split(",")):
    os.environ["FEMB200_ASM_CH"], os.environ["FEMB200_ASM_TPN"] = ch, tpn
    try:
        t = timeit(lambda: fem.assemble_matrix(A, form))
    except Exception as ex:
        print("CH", ch, "R", R, "failed:", ex)
        continue
    chk = A.values.double().square().sum().item()
    if ref is None:
        ref = chk
    print(f"assembly TPN={tpn} CH={ch} R={R}: {t:.3f} ms  {m.ndofs / t / 1e6:.2f} GDOF/s  frac {(8*A.nnz + 32*m.ncells + 16*m.nnodes)/t/1e6/6451.2:.3f} chk {abs(chk-ref)/abs(ref):.1e}", flush=True)
v = torch.randn(m.ndofs, dtype=torch.float64, device="cuda")
y = torch.empty_like(v)
t = timeit(lambda: A.mult(v, y), 20)
print(f"spmv: {t:.3f} ms frac {(36*A.nnz_blocks + 40*m.nnodes)/t/1e6/6451.2:.3f}")
for cfg in os.environ.get("SPMV_CFGS", "6438,6428,6448,3228,3238,3248,6424,6434,6444").split(","):
    os.environ["FEMB200_SPMV_CFG"] = cfg
    try:
        t = timeit(lambda: A.mult(v, y), 20)
        print(f"spmv cfg {cfg} (R,S,L): {t:.3f} ms frac {(36*A.nnz_blocks + 40*m.nnodes)/t/1e6/6451.2:.3f} |y| {y.norm().item():.12e}", flush=True)
    except Exception as ex:
        print("spmv cfg", cfg, "failed", ex)
os.environ.pop("FEMB200_SPMV_CFG", None)
