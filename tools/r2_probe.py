"""Developer probe (round 2): times the kernels under work on one GPU.  python tools/r2_probe.py [what ...]
what: tab (tabulate P2), dmg (damaged reassembly n=1448), asm (linear assembly n=1448), pa (Q2 apply n=2048)"""
import os
import sys
import json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fem-libraries_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from femb200 import fem, dist, mesh as fm

PEAK = 6451.2


def timed(fn, k=10, w=3):
    for _ in range(w):
        fn()
        torch.cuda.synchronize()  # the damaged schedule follows the previous assembly's damage share (host-mapped word)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = [a.elapsed_time(b) for a, b in evs]
    return float(np.mean(t)), float(np.min(t))


def main():
    what = sys.argv[1:] or ["tab", "dmg", "asm"]
    out = {}
    n = 1448
    p = dist.strip_partition_device(n, n, 0, 1)
    m = p.mesh
    form = fem.ElasticityForm(m, p.E, 0.3)
    if "tab" in what:
        Ae = torch.empty((m.ncells, 12, 12), dtype=torch.float64, device="cuda")
        for lay in (0, 1):
            ms, mn = timed(lambda: fem.tabulate_tensor_batched(form, layout=lay, out=Ae))
            out[f"tabulate_p2_layout{lay}"] = {"ms": ms, "min": mn, "frac_hbm": 1192 * m.ncells / (ms * 1e-3) / 1e9 / PEAK,
                                               "tflops": 4314 * m.ncells / (ms * 1e-3) / 1e12}
        dfull = torch.full((m.nnodes,), 0.5, dtype=torch.float64, device="cuda")
        ufull = 1e-3 * torch.randn(2 * m.nnodes, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
        for var, nm in ((0, "closed"), (1, "ad")):
            fd = fem.ElasticityForm(m, p.E, 0.3, d=dfull, u=ufull, variant=var)
            ms, mn = timed(lambda: fem.tabulate_tensor_batched(fd, layout=0, out=Ae))
            out[f"tabulate_p2_damaged_{nm}"] = {"ms": ms, "min": mn}
        del Ae
        m1 = fm.jitter(fm.structured_triangles(2896, order=1), 0.2, seed=1234)
        f1 = fem.ElasticityForm(m1, fm.young_per_cell(m1.ncells), 0.3)
        A1 = torch.empty((m1.ncells, 6, 6), dtype=torch.float64, device="cuda")
        for lay in (0, 1):
            ms, mn = timed(lambda: fem.tabulate_tensor_batched(f1, layout=lay, out=A1))
            out[f"tabulate_p1_layout{lay}"] = {"ms": ms, "min": mn, "frac_hbm": 316 * m1.ncells / (ms * 1e-3) / 1e9 / PEAK,
                                               "gelems": m1.ncells / ms / 1e6}
        del A1, f1, m1
    A = fem.create_matrix(form)
    abytes = 8 * A.nnz + m.ncells * 32 + m.nnodes * 16
    if "asm" in what:
        for so in (1, 0):
            A.set_option("stream_out", so)
            ms, mn = timed(lambda: fem.assemble_matrix(A, form))
            out[f"assemble_p2_stream_out{so}" + ("b" if f"assemble_p2_stream_out{so}" in out else "")] = {"ms": ms, "min": mn, "frac": abytes / (ms * 1e-3) / 1e9 / PEAK}
    if "spmv" in what:
        fem.assemble_matrix(A, form)
        v = torch.randn(2 * m.nnodes, dtype=torch.float64, device="cuda")
        y = torch.empty_like(v)
        sb = 34 * A.nnz_blocks + 40 * m.nnodes
        ms, mn = timed(lambda: fem.capi.call("femb200_spmv", A.plan, fem._p(A.values), fem._p(v), fem._p(y), fem._stream()), k=20, w=3)
        out["spmv_p2_n1448"] = {"ms": ms, "min": mn, "frac_moved": sb / (ms * 1e-3) / 1e9 / PEAK}
    if "vec" in what:
        u0 = 1e-3 * torch.randn(2 * m.nnodes, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
        fl = torch.ones(2 * m.nnodes, dtype=torch.float64, device="cuda")
        bout = torch.empty(2 * m.nnodes, dtype=torch.float64, device="cuda")
        dband = torch.clamp(1.0 - (m.x[:, 1] - 0.5 - 0.1 * torch.sin(6.0 * m.x[:, 0])).abs() / 10.0, min=0.0, max=0.95)
        for label, dd, ff in (("linear", None, None), ("linear_load", None, fl), ("damaged100_load", dband, fl)):
            fv = fem.ElasticityForm(m, p.E, 0.3, d=dd, u=u0)
            ms, mn = timed(lambda: fem.assemble_vector(A, fv, ff, out=bout), k=5, w=2)
            out[f"vector_{label}"] = {"ms": ms, "min": mn}
    if "dmg" in what or "dmg100" in what:
        xy = m.x
        u = 1e-3 * torch.randn(2 * m.nnodes, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
        levels = (("100pct", 10.0),) if "dmg100" in what else (("10pct", 0.05), ("50pct", 0.25), ("100pct", 10.0))
        for label, hw in levels:
            d = torch.clamp(1.0 - (xy[:, 1] - 0.5 - 0.1 * torch.sin(6.0 * xy[:, 0])).abs() / hw, min=0.0, max=0.95)
            for variant, name in ((0, "closed"), (1, "ad")):
                f2 = fem.ElasticityForm(m, p.E, 0.3, d=d, u=u, variant=variant)
                for stg in ((0, 1, 2) if "stage" in what else (0,)):
                    A.set_option("damage_stage", stg)
                    ms, mn = timed(lambda: fem.assemble_matrix(A, f2), k=5, w=2)
                    out[f"dmg_{label}_{name}" + (f"_stage{stg}" if "stage" in what else "")] = {"ms": ms, "frac": abytes / (ms * 1e-3) / 1e9 / PEAK}
    if "p1" in what:
        n1 = 2896
        m1 = fm.jitter(fm.structured_triangles(n1, order=1), 0.2, seed=1234)
        f1 = fem.ElasticityForm(m1, fm.young_per_cell(m1.ncells), 0.3)
        A1 = fem.create_matrix(f1)
        b1 = 8 * A1.nnz + m1.ncells * (3 * 4 + 8) + m1.nnodes * 16
        for so in (1, 0):
            A1.set_option("stream_out", so)
            ms, mn = timed(lambda: fem.assemble_matrix(A1, f1))
            out[f"assemble_p1_n{n1}_stream_out{so}" + ("b" if f"assemble_p1_n{n1}_stream_out{so}" in out else "")] = {"ms": ms, "min": mn, "frac": b1 / (ms * 1e-3) / 1e9 / PEAK,
                                                        "gdofs": m1.ndofs / ms / 1e6, "nnz": A1.nnz}
        del A1, f1, m1
    if "pa" in what:
        del A
        nq = 2048
        mq = fm.jitter(fm.structured_quads_q2(nq), 0.2, seed=1234)
        Eq = fm.young_per_cell(mq.ncells)
        bc, g = fm.dirichlet_markers(mq)
        fq = fem.ElasticityForm(mq, Eq, 0.3)
        pa = fem.PAOperator(fq, bcs=[fem.DirichletBC(bc, g)])
        v = torch.randn(mq.ndofs, dtype=torch.float64, device="cuda")
        y = torch.empty_like(v)
        ms, mn = timed(lambda: pa.mult(v, y))
        nbytes = mq.nnodes * 32 + mq.ncells * (8 * (2 * mq.nv + 2) + 4 * mq.nd + 4)
        out["pa_q2_n2048"] = {"ms": ms, "min": mn, "frac": nbytes / (ms * 1e-3) / 1e9 / PEAK, "gdofs": mq.ndofs / ms / 1e6}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
