"""Multi-GPU path on the device.

* one GPU: the strips of 2 and 3 ranks are built one after the other on the same
  device; the halo is applied by hand (slice copies), the kernels run with their
  owned-row ranges: the concatenated owned CSR rows must equal the single-GPU CSR
  bit for bit, values included, and the distributed SpMV the global one;
* >= 2 GPUs (gpurun --gpus 2): real NCCL ranks, DistCG against the oracle PCG.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle  # noqa: E402
from femb200 import mesh as fm  # noqa: E402

NX, ROWS = 24, 10


def global_mesh(world):
    from femb200 import dist
    m = fm.structured_triangles(NX, ROWS * world, order=2, ly=ROWS * world / NX)
    return dist.jitter_rows(m, 0.2, 1234, 0, 2 * ROWS * world + 1)


@pytest.mark.parametrize("world", [2, 3])
def test_strips_match_global_on_one_gpu(world):
    import torch
    from femb200 import dist, fem
    mg = global_mesh(world)
    Eg = fm.young_per_cell(mg.ncells)
    fg = fem.ElasticityForm(mg, Eg)
    Ag = fem.assemble_matrix(fem.create_matrix(fg), fg)
    rp_g, ci_g, va_g = Ag.rowptr.cpu().numpy(), Ag.colidx.cpu().numpy(), Ag.values.cpu().numpy()
    vg = torch.randn(mg.ndofs, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    want = Ag.mult(vg).cpu().numpy()
    got = np.full(mg.ndofs, np.nan)
    dots = 0.0
    for r in range(world):
        p = dist.strip_partition(NX, ROWS * world, 2, r, world)
        f = fem.ElasticityForm(p.mesh, p.E)
        A = fem.assemble_matrix(fem.create_matrix(f), f)
        rp, ci, va = A.rowptr.cpu().numpy(), A.colidx.cpu().numpy(), A.values.cpu().numpy()
        lo, hi, go = 2 * p.own_lo, 2 * p.own_hi, 2 * (p.node_offset + p.own_lo)
        # global CSR identity: structure bit-exact, values bit-identical
        np.testing.assert_array_equal(np.diff(rp[lo:hi + 1]), np.diff(rp_g[go:go + hi - lo + 1]))
        seg, gseg = slice(rp[lo], rp[hi]), slice(rp_g[go], rp_g[go + hi - lo])
        np.testing.assert_array_equal(ci[seg] + 2 * p.node_offset, ci_g[gseg])
        np.testing.assert_array_equal(va[seg], va_g[gseg])
        # halo by hand: the local vector is a window of the global one
        v = vg[2 * p.node_offset:2 * (p.node_offset + p.mesh.nnodes)].clone()
        y = torch.full_like(v, float("nan"))
        A.set_row_range(p.own_lo, p.own_hi)
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        fem.capi.call("femb200_spmv_dot", A.plan, fem._p(A.values), fem._p(v), fem._p(y), fem._p(out), fem._stream())
        yh = y.cpu().numpy()
        assert np.isnan(yh[:lo]).all() and np.isnan(yh[hi:]).all()
        got[go:go + hi - lo] = yh[lo:hi]
        dots += out.item()
    np.testing.assert_array_equal(got, want)
    ref = (vg.cpu().numpy() * want).sum()
    assert abs(dots - ref) <= 1e-12 * np.abs(vg.cpu().numpy() * want).sum()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, out):
    import torch
    import torch.distributed as td
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from femb200 import dist, fem
        p = dist.strip_partition(NX, ROWS * world, 2, rank, world)
        f = fem.ElasticityForm(p.mesh, p.E)
        A = fem.create_matrix(f)
        fem.assemble_matrix(A, f, bcs=[fem.DirichletBC(p.bc, p.g)])
        # right-hand side of the global lifted problem, computed with the oracle on every rank
        mg = global_mesh(world)
        Eg = fm.young_per_cell(mg.ncells)
        bcg, gg = fm.dirichlet_markers(mg)
        rowptr, colidx = oracle.build_pattern(mg.nnodes, mg.dofmap)
        full = oracle.assemble_matrix(mg.etype, mg.x, mg.xdofmap, mg.dofmap, Eg, 0.3, rowptr, colidx)
        vals = oracle.assemble_matrix(mg.etype, mg.x, mg.xdofmap, mg.dofmap, Eg, 0.3, rowptr, colidx, bc=bcg)
        b = -oracle.spmv(rowptr, colidx, full, gg)
        b[bcg != 0] = gg[bcg != 0]
        bl = fem.to_device(b[2 * p.node_offset:2 * (p.node_offset + p.mesh.nnodes)], np.float64)
        x = torch.zeros_like(bl)
        cg0 = dist.DistCG(A, p, rel_tol=1e-12, max_iter=4000)            # plain: halo, then all owned rows
        x0 = torch.zeros_like(bl)
        cg0.solve(bl, x0)
        cg = dist.DistCG(A, p, rel_tol=1e-12, max_iter=4000, overlap=True)
        # the overlapped operator apply (interior rows during the halo exchange, then the rows next to the
        # ghosts) against the oracle's global product on the owned rows
        vg = np.random.default_rng(3).standard_normal(mg.ndofs)
        vl = fem.to_device(vg[2 * p.node_offset:2 * (p.node_offset + p.mesh.nnodes)].copy(), np.float64)
        vl[:2 * p.own_lo] = float("nan")
        vl[2 * p.own_hi:] = float("nan")                      # ghosts must come from the neighbours
        yl = torch.full_like(vl, float("nan"))
        cg.mult(vl, yl)
        torch.cuda.synchronize()
        wantl = oracle.spmv(rowptr, colidx, vals, vg)[2 * (p.node_offset + p.own_lo):2 * (p.node_offset + p.own_hi)]
        gotl = yl[2 * p.own_lo:2 * p.own_hi].cpu().numpy()
        assert np.isfinite(gotl).all() and np.linalg.norm(gotl - wantl) <= 1e-12 * np.linalg.norm(wantl)
        cg.solve(bl, x)
        assert cg0.converged and abs(cg0.iterations - cg.iterations) <= 1
        assert ((x0 - x)[2 * p.own_lo:2 * p.own_hi].norm() / x[2 * p.own_lo:2 * p.own_hi].norm()).item() < 1e-10
        xs = dist.gather_owned(p, x)
        if rank == 0:
            want, it, _, conv = oracle.pcg(rowptr, colidx, vals, b, rtol=1e-12, maxit=4000, jacobi=True)
            err = np.linalg.norm(xs - want) / np.linalg.norm(want)
            assert cg.converged and conv and err < 1e-10, f"err {err} its {cg.iterations} vs {it}"
            assert abs(cg.iterations - it) <= max(3, it // 50)
        out.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        out.put((rank, traceback.format_exc()))
        raise
    finally:
        td.destroy_process_group()


def test_dist_cg_nccl():
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [out.get(timeout=300) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"
