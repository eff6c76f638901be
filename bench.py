#!/usr/bin/env python
"""bench.py -- the hot path of the mechanic2d elasticity examples on N B200s.

Workload (BASELINE.json configs[1]): P2 triangles, structured n x n cells split
by the right diagonal (n = 1448 -> 4 193 408 elements, 16 785 218 dofs, 385 886 212
CSR non-zeros per GPU), jittered vertices, the reference's 200-value Young-modulus
table, nu = 0.3, Dirichlet x = 0 / x = 1.  At N > 1 every rank owns an n x n strip of
a [0,1] x [0,N] domain (weak scaling), assembles it without communication (one
ghost row of cells) and runs CG with NCCL halo exchange + all-reduce.

One timed STEP = one full matrix assembly (element integration + scatter into
the CSR + Dirichlet rows/cols), the setJ lambda of the reference (F.cc:847-862).
`value` = dofs assembled per second over all ranks (GDOF/s).  The same run also
times the operator apply inside CG (SpMV alone and whole PCG iterations) and
reports them, with their own roofline, under "cg".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "fem-libraries_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "assembly GDOF/s (P2 elasticity, CSR) with CG SpMV GB/s alongside"
UNIT = "GDOF/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------
# algorithmic bytes
# ---------------------------------------------------------------------------
def assembly_bytes(nnz: int, ncells: int, nnodes: int) -> int:
    """write nnz values once + read connectivity (6 x int32), E (8 B) per cell and
    coordinates (16 B) per node: ~808 B per P2 element."""
    return 8 * nnz + ncells * (6 * 4 + 8) + nnodes * 16


def spmv_bytes(nnz_blocks: int, nnodes: int) -> int:
    """block pattern: 32 B of values + 4 B column index per 2x2 block; per node 8 B of
    row pointer, 16 B of x read, 16 B of y written."""
    return 36 * nnz_blocks + nnodes * (8 + 16 + 16)


def cg_iter_bytes(nnz_blocks: int, nnodes: int, jacobi: bool = True) -> int:
    """SpMV + update_xr (read d, Ad, x, r, dinv; write x, r) + update_dir (read r, dinv,
    d; write d), 16 B per node per vector pass."""
    passes = (7 if jacobi else 6) + (4 if jacobi else 3)
    return spmv_bytes(nnz_blocks, nnodes) + passes * 16 * nnodes


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])), mx.append(float(r[1])), pw.append(float(r[2]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if r[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def build_problem(n: int, rank: int, world: int):
    """Rank-local mesh (with its ghost layers at N > 1), materials and Dirichlet data."""
    from femb200 import mesh as fm
    if world == 1:
        m = fm.jitter(fm.structured_triangles(n, order=2), 0.2, seed=1234)
        E = fm.young_per_cell(m.ncells)
        bc, g = fm.dirichlet_markers(m)
        return m, E, bc, g, None
    from femb200 import dist
    part = dist.strip_partition(n, n * world, order=2, rank=rank, world=world, jitter_amp=0.2, seed=1234)
    return part.mesh, part.E, part.bc, part.g, part


# ---------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference (oracle) on the host cores
# ---------------------------------------------------------------------------
def cpu_reference(n_sample: int, steps: int, warmup: int, spmv_reps: int = 5):
    from oracle import oracle
    from femb200 import mesh as fm
    m = fm.jitter(fm.structured_triangles(n_sample, order=2), 0.2, seed=1234)
    E = fm.young_per_cell(m.ncells)
    bc, _ = fm.dirichlet_markers(m)
    rowptr, colidx = oracle.build_pattern(m.nnodes, m.dofmap)
    vals = np.empty(int(rowptr[-1]))
    nt_all = oracle.num_threads()
    best = None
    for nt in sorted({1, nt_all}):
        ts = []
        for i in range(warmup + steps):
            t = time.perf_counter()
            oracle.assemble_matrix(m.etype, m.x, m.xdofmap, m.dofmap, E, 0.3, rowptr, colidx, bc=bc, nthreads=nt,
                                   values=vals)
            if i >= warmup:
                ts.append(time.perf_counter() - t)
        t_asm = float(np.mean(ts))
        v = np.random.default_rng(0).standard_normal(m.ndofs)
        y = np.empty(m.ndofs)
        oracle.spmv(rowptr, colidx, vals, v, nthreads=nt, y=y)
        t = time.perf_counter()
        for _ in range(spmv_reps):
            oracle.spmv(rowptr, colidx, vals, v, nthreads=nt, y=y)
        t_spmv = (time.perf_counter() - t) / spmv_reps
        rec = {"threads": nt, "assembly_s": t_asm, "assembly_gdofs": m.ndofs / t_asm / 1e9, "spmv_s": t_spmv,
               "spmv_gbs": (12 * int(rowptr[-1]) + 24 * m.ndofs + 8) / t_spmv / 1e9,
               "spmv_gdofs": m.ndofs / t_spmv / 1e9}
        if best is None or rec["assembly_gdofs"] > best["assembly_gdofs"]:
            best = rec
    best["sample"] = (f"P2 n={n_sample} ({m.ncells} elements, {m.ndofs} dofs) of the n=1448 workload, "
                      f"{steps} assemblies after {warmup} warm-ups, oracle/fem_oracle.c -O3 OpenMP")
    best["ms_per_step"] = 1e3 * best["assembly_s"]
    return best


def reference_element_kernel_rate(n_elems: int = 400000):
    """The reference's OWN element code (damIntegrator::AssembleElementGrad, M.cc:639-916, compiled in place
    into oracle/_ref by oracle/ref_shim/build_ref.sh), one host thread, P1 triangles, d = 0: elements per
    second, next to the 5.5 M elements/s/core the reference publishes (curve_time.txt col 84).  None when
    oracle/_ref does not exist on this box."""
    import ctypes as C
    so = os.path.join(ROOT, "oracle", "_ref", "libref_B.so")
    if not os.path.exists(so):
        return None
    L = C.CDLL(so)
    if not hasattr(L, "ref_element_grad_batch"):
        return None
    dp = C.POINTER(C.c_double)
    L.ref_element_grad_batch.argtypes = [C.c_long, dp, dp, dp, dp, dp]
    rng = np.random.default_rng(0)
    xv = np.tile(np.array([0.0, 0.0, 1.0, 0.1, 0.2, 0.9]), (n_elems, 1)) + 0.05 * rng.random((n_elems, 6))
    lam, mu, d = np.full(n_elems, 4.0e7), np.full(n_elems, 2.7e7), np.zeros(n_elems)
    out = np.empty((n_elems, 36))
    P = lambda a: a.ctypes.data_as(dp)
    L.ref_element_grad_batch(1000, P(xv), P(lam), P(mu), P(d), P(out))
    t = time.perf_counter()
    L.ref_element_grad_batch(n_elems, P(xv), P(lam), P(mu), P(d), P(out))
    dt = time.perf_counter() - t
    return {"melems_per_s_per_core": n_elems / dt / 1e6, "elements": n_elems,
            "what": "the reference's own damIntegrator::AssembleElementGrad (M.cc:639-916, P1, d = 0) compiled in place "
                    "against the MFEM stand-in of oracle/ref_shim, 1 thread, -O3 -DNDEBUG",
            "published_melems_per_s_per_core": 5.5}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(args.cpu_n, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["assembly_gdofs"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "P2 triangles, structured n=1448 (4 193 408 elements), assembly + CG SpMV",
                       "timed_on": r["sample"]},
            "cpu_baseline": {"value": r["assembly_gdofs"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                             "sample": r["sample"]},
            "cg": {"spmv_gbs": r["spmv_gbs"], "spmv_gdofs": r["spmv_gdofs"]},
            "e2e": {"value": r["assembly_gdofs"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as td
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    from femb200 import fem
    K, W, n = args.steps, max(args.warmup, 0), args.n

    m, E, bc, g, part = build_problem(n, rank, world)
    form = fem.ElasticityForm(m, E, 0.3)
    # load the library's CUDA module (lazy, one-off) outside the pattern-build timing
    from femb200 import mesh as _fm
    _tiny = _fm.structured_triangles(4, order=2)
    fem.create_matrix(fem.ElasticityForm(_tiny, _fm.young_per_cell(_tiny.ncells), 0.3))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    A = fem.create_matrix(form)
    A.set_bcs([fem.DirichletBC(bc, g)])
    torch.cuda.synchronize()
    pattern_ms = 1e3 * (time.perf_counter() - t0)
    owned_nodes = m.nnodes if part is None else part.n_owned
    owned_dofs = 2 * owned_nodes
    owned_cells = m.ncells if part is None else part.n_owned_cells

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def maxtime(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    def sumint(v: int) -> int:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.int64, device="cuda")
        td.all_reduce(t)
        return int(t.item())

    def timed(fn, k):
        """k calls bracketed by barrier + synchronize; device time by CUDA events on the
        launching stream; returns (total ms max over ranks, per-call ms list of this rank)."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for a, b in evs:
            a.record()
            fn()
            b.record()
        e.record()
        barrier()
        return maxtime(s.elapsed_time(e)), [a.elapsed_time(b) for a, b in evs]

    # ---- assembly: the timed step -----------------------------------------
    def step():
        fem.assemble_matrix(A, form)

    for _ in range(W):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms, per_call = timed(step, K)
    asm_ms = total_ms / K
    kernel_ms = float(np.mean(per_call))
    total_dofs = sumint(owned_dofs)
    total_cells = sumint(owned_cells)
    value = total_dofs / (asm_ms * 1e-3) / 1e9

    # ---- operator apply inside CG ------------------------------------------
    if part is None:
        bvec = fem.to_device(np.where(bc != 0, g, 1.0), np.float64)
        cg = fem.CGSolver(rel_tol=0.0, abs_tol=0.0, max_iter=args.cg_iters)
        cg.SetOperator(A)
        cg.SetPreconditioner("jacobi")
        xsol = torch.empty_like(bvec)
        ytmp = torch.empty_like(bvec)

        def cg_run():
            cg.Mult(bvec, xsol, fixed_iters=args.cg_iters)

        def spmv_run():
            A.mult(bvec, ytmp)
    else:
        from femb200 import dist
        dcg = dist.DistCG(A, part, rel_tol=0.0, abs_tol=0.0, max_iter=args.cg_iters)
        bvec = fem.to_device(np.where(bc != 0, g, 1.0), np.float64)
        xsol = torch.zeros_like(bvec)
        ytmp = torch.empty_like(bvec)

        def cg_run():
            dcg.solve(bvec, xsol, fixed_iters=args.cg_iters)

        def spmv_run():
            dcg.mult(bvec, ytmp)     # halo exchange, then the owned rows

    for _ in range(max(1, min(W, 2))):
        cg_run()
        spmv_run()
    kc = max(1, min(K, 5))
    cg_total, _ = timed(cg_run, kc)
    # one PCG call = setup (init + first apply) + cg_iters iterations: count cg_iters + 1 applies
    cg_iter_ms = cg_total / kc / (args.cg_iters + 1)
    ks = max(K, 10)
    spmv_total, spmv_calls = timed(spmv_run, ks)
    spmv_ms = spmv_total / ks
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: host buffers through the public API -------------------
    gv = form.geometry_vertices                         # the geometry is P1: only its vertices are inputs
    hx = torch.from_numpy(np.ascontiguousarray(m.x[gv])).pin_memory()
    dxv = torch.empty(hx.shape, dtype=torch.float64, device="cuda")
    hE = torch.from_numpy(np.ascontiguousarray(E)).pin_memory()
    # Two steps in flight: a copy stream uploads the inputs of step k + 1 (double-buffered device
    # inputs) while the compute stream assembles step k; every step's H2D copies, its assembly, its
    # checksum and the D2H read of that checksum are inside the timed region, and the host reads the
    # result of step k - 1 before it enqueues step k + 1.
    copy_stream = torch.cuda.Stream()
    comp = torch.cuda.current_stream()
    dxv = [torch.empty(hx.shape, dtype=torch.float64, device="cuda") for _ in range(2)]
    dE = [torch.empty(hE.shape, dtype=torch.float64, device="cuda") for _ in range(2)]
    hout = [torch.empty(2, dtype=torch.float64).pin_memory() for _ in range(2)]
    dout = [torch.empty(2, dtype=torch.float64, device="cuda") for _ in range(2)]
    ev_up = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    state = {"k": 0}

    def e2e_step():
        k = state["k"]
        b = k & 1
        with torch.cuda.stream(copy_stream):
            if k >= 2:
                copy_stream.wait_event(ev_done[b])         # the buffers of step k - 2 are free
            dxv[b].copy_(hx, non_blocking=True)            # H2D: vertex coordinates of this step
            dE[b].copy_(hE, non_blocking=True)             # H2D: material field of this step
            ev_up[b].record(copy_stream)
        comp.wait_event(ev_up[b])
        form.set_geometry(dxv[b])
        form.set_E(dE[b])
        fem.assemble_matrix(A, form, norms_out=dout[b])   # assembly + Dirichlet, (|K|_F^2, trace K) fused in
        hout[b].copy_(dout[b], non_blocking=True)          # D2H: (|K|_F^2, trace K)
        ev_done[b].record(comp)
        if k >= 1:
            ev_done[1 - b].synchronize()                   # the host consumes the result of step k - 1
            state["last"] = (float(hout[1 - b][0]), float(hout[1 - b][1]))
        state["k"] = k + 1

    def e2e_drain():
        torch.cuda.synchronize()
        b = (state["k"] - 1) & 1
        state["last"] = (float(hout[b][0]), float(hout[b][1]))

    for _ in range(max(1, min(W, 2))):
        e2e_step()
    e2e_drain()

    def e2e_all():
        for _ in range(K):
            e2e_step()
        e2e_drain()

    e2e_total, _ = timed(e2e_all, 1)
    e2e_ms = e2e_total / K
    e2e_value = total_dofs / (e2e_ms * 1e-3) / 1e9
    h2d = hx.numel() * 8 + hE.numel() * 8
    fro, tr = float(np.sqrt(state["last"][0])), float(state["last"][1])

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    a_bytes = assembly_bytes(A.nnz, m.ncells, m.nnodes)
    s_bytes = spmv_bytes(A.nnz_blocks if part is None else part.owned_nnz_blocks(A), owned_nodes)
    c_bytes = cg_iter_bytes(A.nnz_blocks if part is None else part.owned_nnz_blocks(A), owned_nodes)
    a_gbs = a_bytes / (kernel_ms * 1e-3) / 1e9
    s_gbs = s_bytes / (float(np.mean(spmv_calls)) * 1e-3) / 1e9
    c_gbs = c_bytes / (cg_iter_ms * 1e-3) / 1e9

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": asm_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"P2 triangles, structured n={n} per GPU ({m.ncells} local elements incl. ghost row), "
                               "jittered, E = reference 200-value table, nu = 0.3; step = full CSR assembly + Dirichlet",
                   "elements": total_cells, "dofs": total_dofs, "nnz_per_gpu": A.nnz,
                   "l2": "inputs + outputs per step (3.5 GB) exceed the 126 MB L2; no explicit flush",
                   "parallelism": f"strips x{world}" if world > 1 else "single GPU", "cg_iters": args.cg_iters,
                   "pattern_build_ms": pattern_ms},
        "roofline": {"kernel": "assemble_fast_kernel<P2> (+ cell_setup_kernel, dirichlet_kernel)", "bound": "hbm", "achieved": a_gbs,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": a_gbs / peak,
                     "algorithmic_bytes_per_launch": a_bytes, "kernel_ms": kernel_ms,
                     "traffic": traffic_from_profile("assemble") if (n == 1448 and world == 1) else None},
        "cg": {"spmv_ms": spmv_ms, "spmv_gbs": s_gbs, "spmv_frac": s_gbs / peak, "spmv_gdofs": owned_dofs * world / (spmv_ms * 1e-3) / 1e9,
               "spmv_algorithmic_bytes": s_bytes, "spmv_traffic": traffic_from_profile("spmv") if (n == 1448 and world == 1) else None,
               "cg_iter_ms": cg_iter_ms, "cg_iter_gbs": c_gbs, "cg_iter_frac": c_gbs / peak,
               "cg_iter_gdofs": total_dofs / (cg_iter_ms * 1e-3) / 1e9, "precond": "jacobi", "iters": args.cg_iters},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 16, "what": "per step: pinned host vertex coordinates + E -> device (copy stream, "
                                                  "double-buffered, overlapping the previous step's assembly), set_geometry, "
                                                  "assemble_matrix(A, form, bcs, norms_out) (Frobenius norm and trace fused into the assembly pass), 16-byte read back consumed by "
                                                  "the host; K steps timed as a whole", "fro": fro, "trace": tr},
        "gpu_launches": 3 * K,  # cell_setup + assemble + dirichlet per step
        "clocks": clocks,
    }
    if world == 1:
        line["matrix_free"] = pa_probe(fem, form, bc, g, m, timed, peak)
    if world == 1 and args.extras:
        line["extras"] = extras(fem, timed, peak)
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(args.cpu_n, 3, 1)
        line["cpu_baseline"] = {"value": r["assembly_gdofs"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                                "sample": r["sample"], "spmv_gbs": r["spmv_gbs"], "spmv_gdofs": r["spmv_gdofs"],
                                "reference_element_kernel": reference_element_kernel_rate()}
    print(json.dumps(line), flush=True)
    if world > 1:
        td.destroy_process_group()


def pa_probe(fem, form, bc, g, m, timed, peak):
    """Matrix-free apply (AssemblePA / AddMultPA role) of the same operator on the same mesh."""
    import torch
    pa = fem.PAOperator(form, bcs=[fem.DirichletBC(bc, g)])
    v = torch.randn(m.ndofs, dtype=torch.float64, device="cuda")
    y = torch.empty_like(v)
    for _ in range(3):
        pa.mult(v, y)
    tot, _ = timed(lambda: pa.mult(v, y), 10)
    ms = tot / 10
    nbytes = m.nnodes * 32 + m.ncells * (8 * (2 * m.nv + 2) + 4 * m.nd + 4)
    return {"pa_apply_ms": ms, "gdofs": m.ndofs / (ms * 1e-3) / 1e9, "gbs": nbytes / (ms * 1e-3) / 1e9,
            "frac": nbytes / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": nbytes}


def extras(fem, timed, peak):
    """Other BASELINE configs on one GPU (developer flag --extras): config 3 (Q2 quads, matrix-free,
    16.8 M elements) and the config-5 workload (damaged-tangent reassembly, closed form and AD)."""
    import torch
    from femb200 import mesh as fm
    out = {}
    n = 4096
    m = fm.jitter(fm.structured_quads_q2(n), 0.2, seed=1234)
    E = fm.young_per_cell(m.ncells)
    bc, g = fm.dirichlet_markers(m)
    form = fem.ElasticityForm(m, E, 0.3)
    r = pa_probe(fem, form, bc, g, m, timed, peak)
    pa = fem.PAOperator(form, bcs=[fem.DirichletBC(bc, g)])
    cg = fem.CGSolver(rel_tol=0.0, max_iter=10)
    cg.SetOperator(pa)
    cg.SetPreconditioner("jacobi")
    b = torch.ones(m.ndofs, dtype=torch.float64, device="cuda")
    x = torch.empty_like(b)
    cg.Mult(b, x, fixed_iters=10)
    tot, _ = timed(lambda: cg.Mult(b, x, fixed_iters=10), 3)
    r.update({"workload": f"Q2 quads n={n}: {m.ncells} elements, {m.ndofs} dofs, 3x3 Gauss, sum-factorised",
              "pa_cg_iter_ms": tot / 3 / 11, "pa_cg_iter_gdofs": m.ndofs / (tot / 3 / 11 * 1e-3) / 1e9,
              "survey_bytes_per_element": 596, "frac_at_survey_bytes": 596 * m.ncells / (r["pa_apply_ms"] * 1e-3) / 1e9 / peak})
    out["config3_q2_matrix_free"] = r
    del pa, cg, form, b, x
    torch.cuda.empty_cache()
    n = 1448
    m = fm.jitter(fm.structured_triangles(n, order=2), 0.2, seed=1234)
    E = fm.young_per_cell(m.ncells)
    d = fm.damage_band(m)
    u = 1e-3 * np.random.default_rng(0).standard_normal(m.ndofs)
    res = {"workload": f"P2 n={n}, damage band on {100 * float((d[m.xdofmap].mean(axis=1) > 0).mean()):.1f} % of the cells, "
                       "values-only reassembly on a frozen pattern"}
    for variant, name in ((0, "closed_form"), (1, "ad")):
        form = fem.ElasticityForm(m, E, 0.3, d=d, u=u, variant=variant)
        A = fem.create_matrix(form)
        for _ in range(2):
            fem.assemble_matrix(A, form)
        tot, _ = timed(lambda: fem.assemble_matrix(A, form), 5)
        res[name + "_ms"] = tot / 5
        res[name + "_gdofs"] = m.ndofs / (tot / 5 * 1e-3) / 1e9
        del A, form
    out["config5_damaged_reassembly_1gpu"] = res
    return out


def config5(args, rank, world, local):
    """BASELINE config 5: repeated tangent reassembly (closed-form and AD tangents, M.cc:736-872 / 752-765) of
    16 773 632 P2 elements on `world` GPUs: strips of the nx = 2896 mesh, a vertical damage band through
    every strip, 10 reassemblies on the frozen pattern.  Assembly needs no communication (ghost cell row)."""
    import torch
    import torch.distributed as td
    from femb200 import fem, dist, mesh as fm
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx = 2896
    if world == 1:
        m = fm.jitter(fm.structured_triangles(nx, order=2), 0.2, seed=1234)
        E, owned_dofs = fm.young_per_cell(m.ncells), m.ndofs
    else:
        part = dist.strip_partition(nx, nx - nx % world, order=2, rank=rank, world=world, jitter_amp=0.2, seed=1234)
        m, E, owned_dofs = part.mesh, part.E, 2 * part.n_owned
    x, y = m.x[:, 0], m.x[:, 1]
    d = np.minimum(np.maximum(0.0, 1.0 - np.abs(x - 0.5 - 0.1 * np.sin(6.0 * y)) / 0.05), 0.95)
    u = 1e-3 * np.random.default_rng(rank).standard_normal(m.ndofs)
    share = float((d[m.xdofmap].mean(axis=1) > 0).mean())
    reps = 10

    def sync():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    out = {}
    for variant, name in ((0, "closed_form"), (1, "ad")):
        form = fem.ElasticityForm(m, E, 0.3, d=d, u=u, variant=variant)
        A = fem.create_matrix(form)
        for _ in range(3):
            fem.assemble_matrix(A, form)
        sync()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fem.assemble_matrix(A, form)
        e.record()
        sync()
        t = torch.tensor([s.elapsed_time(e), float(owned_dofs)], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax = t.clone()
            td.all_reduce(tmax, op=td.ReduceOp.MAX)
            td.all_reduce(t, op=td.ReduceOp.SUM)
            ms, dofs = tmax[0].item() / reps, t[1].item()
        else:
            ms, dofs = t[0].item() / reps, t[1].item()
        out[name] = {"ms_per_reassembly": ms, "gdofs": dofs / (ms * 1e-3) / 1e9}
        del A, form
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"metric": "damaged-tangent reassembly GDOF/s (BASELINE config 5)", "unit": UNIT,
                          "value": out["closed_form"]["gdofs"], "n_gpus": world, "reassemblies": reps,
                          "higher_is_better": True, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": f"P2 triangles nx = {nx}, strips over {world} GPU(s), vertical damage band "
                                                 f"on {100 * share:.1f} % of the cells, values-only reassembly on a frozen "
                                                 "pattern", "elements_total": 2 * nx * (nx - nx % world)},
                          "closed_form": out["closed_form"], "ad": out["ad"],
                          "ad_over_closed": out["ad"]["ms_per_reassembly"] / out["closed_form"]["ms_per_reassembly"]}),
              flush=True)
    if world > 1:
        td.destroy_process_group()


def traffic_from_profile(which: str):
    """dram bytes per launch from the committed ncu --set full summary, if present."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(which)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=1448, help="cells per side per GPU (1448 -> 4.19 M P2 triangles)")
    ap.add_argument("--cg-iters", type=int, default=25)
    ap.add_argument("--cpu-n", type=int, default=512, help="cells per side of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", action="store_true", help="also time config 3 (Q2 matrix-free) and config 5 (damage)")
    ap.add_argument("--config5", action="store_true",
                    help="BASELINE config 5 instead of the headline step: repeated damaged-tangent reassembly "
                         "(closed form and AD) of 16.8 M P2 elements split over the ranks (nx = 2896)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.config5:
        config5(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                int(os.environ.get("LOCAL_RANK", "0")))
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
