#!/usr/bin/env python
"""Golden (|K|_F^2, trace K) of the CONSTRAINED tangent matrix of the synthetic P2 workloads of bench.py
(BASELINE.json configs[1] n = 1448 and configs[3] n = 5792), computed with the CPU oracle strip by strip
(each strip = one rank of dist.strip_partition, owned rows only, so the sums tile the global matrix) and,
where the global matrix fits in host memory (n = 1448), the state of the oracle's Jacobi-PCG after 25
iterations on the bench right-hand side.

    python tests/golden/make_config_norms.py 1448 5792      ->  tests/golden/config_norms.json

bench.py all-reduces the same sums over its ranks and asserts agreement to 1e-12 (relative) at every N.
Test infrastructure: the product never reads this file."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "fem-libraries_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import oracle  # noqa: E402
from femb200 import dist  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "config_norms.json")
NT = len(os.sched_getaffinity(0))


def strip_sums(n, strips):
    fro2, trace, nnz_owned, ndofs = 0.0, 0.0, 0, 0
    parts = []
    for r in range(strips):
        t = time.time()
        p = dist.strip_partition(n, n, 2, r, strips)
        m = p.mesh
        rowptr, colidx = oracle.build_pattern(m.nnodes, m.dofmap)
        vals = oracle.assemble_matrix(m.etype, m.x, m.xdofmap, m.dofmap, p.E, 0.3, rowptr, colidx, bc=p.bc, nthreads=NT)
        lo, hi = 2 * p.own_lo, 2 * p.own_hi
        seg = slice(int(rowptr[lo]), int(rowptr[hi]))
        v = vals[seg]
        f2 = float(np.sum(v * v))
        rows = np.repeat(np.arange(lo, hi, dtype=np.int32), np.diff(rowptr[lo:hi + 1]))
        tr = float(np.sum(v[colidx[seg] == rows]))
        parts.append({"strip": r, "fro2": f2.hex(), "trace": tr.hex(), "nnz": int(v.size)})
        fro2 += f2
        trace += tr
        nnz_owned += int(v.size)
        ndofs += hi - lo
        print(f"n={n} strip {r + 1}/{strips}: {time.time() - t:.1f} s", flush=True)
        del vals, colidx, rowptr, v, rows
    return {"n": n, "elements": 2 * n * n, "ndofs": ndofs, "nnz": nnz_owned, "fro2": fro2, "trace": trace,
            "fro2_hex": fro2.hex(), "trace_hex": trace.hex(), "strips": strips, "partials": parts}


def pcg_state(n, iters=25):
    """oracle Jacobi-PCG, `iters` iterations from x = 0 on b = g on the Dirichlet dofs, 1 elsewhere."""
    p = dist.strip_partition(n, n, 2, 0, 1)
    m = p.mesh
    rowptr, colidx = oracle.build_pattern(m.nnodes, m.dofmap)
    vals = oracle.assemble_matrix(m.etype, m.x, m.xdofmap, m.dofmap, p.E, 0.3, rowptr, colidx, bc=p.bc, nthreads=NT)
    b = np.where(p.bc != 0, p.g, 1.0)
    x, it, fn, conv = oracle.pcg(rowptr, colidx, vals, b, rtol=0.0, atol=0.0, maxit=iters, jacobi=True, nthreads=NT)
    w = np.cos(np.arange(x.size, dtype=np.float64))
    return {"iters": it, "final_norm": fn, "x_dot_cos": float(np.dot(x, w)), "x_norm": float(np.linalg.norm(x))}


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1448, 5792]
    out = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            out = json.load(f)
    for n in sizes:
        rec = strip_sums(n, 16 if n > 3000 else 4)
        if n <= 1448:
            rec["pcg25"] = pcg_state(n)
        out[str(n)] = rec
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
