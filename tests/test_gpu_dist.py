"""Multi-GPU path on the device.

* one GPU: the strips of 2 and 3 ranks are built one after the other on the same
  device; the halo is applied by hand (slice copies), the kernels run with their
  owned-row ranges: the concatenated owned CSR rows must equal the single-GPU CSR
  bit for bit, values included, and the distributed SpMV the global one;
* >= 2 GPUs (gpurun --gpus 2): real ranks, both transports of csrc/dist.cu (NVLink peer memory and NCCL),
  the distributed operator and DistCG against the oracle's global SpMV / PCG;
* one GPU: the communicator with world = 1 (the path bench.py takes at N = 1) against the oracle.
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
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        A.mult_rows(v, y, p.own_lo, p.own_hi, dot=out)
        yh = y.cpu().numpy()
        assert np.isnan(yh[:lo]).all() and np.isnan(yh[hi:]).all()
        got[go:go + hi - lo] = yh[lo:hi]
        dots += out.item()
    np.testing.assert_array_equal(got, want)
    ref = (vg.cpu().numpy() * want).sum()
    assert abs(dots - ref) <= 1e-12 * np.abs(vg.cpu().numpy() * want).sum()


def assert_norms_match(got, want, noise=1e-7):
    """Residual histories agree to 6 digits; entries at the rounding-noise floor of the converged state only in size."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    big = want > noise
    np.testing.assert_allclose(got[big], want[big], rtol=1e-6)
    assert (got[~big] <= noise).all()


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
        pd = dist.strip_partition_device(NX, ROWS * world, rank, world)      # the generator bench.py uses
        assert np.array_equal(pd.mesh.x.cpu().numpy(), p.mesh.x) and np.array_equal(pd.mesh.dofmap.cpu().numpy(), p.mesh.dofmap)
        assert np.array_equal(pd.E.cpu().numpy(), p.E) and np.array_equal(pd.bc.cpu().numpy(), p.bc)
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
        lo, hi = 2 * p.own_lo, 2 * p.own_hi
        glo = 2 * (p.node_offset + p.own_lo)
        bl = fem.to_device(b[2 * p.node_offset:2 * (p.node_offset + p.mesh.nnodes)], np.float64)
        want, it, _, conv = oracle.pcg(rowptr, colidx, vals, b, rtol=1e-12, maxit=4000, jacobi=True)
        vg = np.random.default_rng(3).standard_normal(mg.ndofs)
        want_mult = oracle.spmv(rowptr, colidx, vals, vg)[glo:glo + hi - lo]
        op = dist.DistOperator(A, p, transport="auto")
        transports = [op.transport] + (["nccl"] if op.transport == "p2p" else [])
        sols = {}
        for tr in transports:
            op.use(tr)
            # the operator apply with ghosts that must come from the neighbours, against the oracle's global product
            vl = fem.to_device(vg[2 * p.node_offset:2 * (p.node_offset + p.mesh.nnodes)].copy(), np.float64)
            vl[:lo] = float("nan")
            vl[hi:] = float("nan")
            yl = torch.full_like(vl, float("nan"))
            op.mult(vl, yl)
            torch.cuda.synchronize()
            gotl = yl[lo:hi].cpu().numpy()
            assert np.isfinite(gotl).all() and np.linalg.norm(gotl - want_mult) <= 1e-12 * np.linalg.norm(want_mult), tr
            # sum over the ranks of a device vector
            s = torch.tensor([rank + 1.0, 0.5, -2.0 * rank], dtype=torch.float64, device="cuda")
            op.allreduce_sum(s)
            assert s.tolist() == [world * (world + 1) / 2, 0.5 * world, -world * (world - 1.0)], (tr, s.tolist())
            for graph in (False, True):
                cg = dist.DistCG(A, p, rel_tol=1e-12, max_iter=4000, op=op, use_graph=graph)
                x = torch.zeros_like(bl)
                cg.solve(bl, x)
                xs = dist.gather_owned(p, x)
                sols[(tr, graph)] = (x[lo:hi].clone(), cg.iterations)
                assert cg.converged, (tr, graph)
                if rank == 0:
                    err = np.linalg.norm(xs - want) / np.linalg.norm(want)
                    assert conv and err < 1e-10, f"{tr} graph={graph}: err {err} its {cg.iterations} vs {it}"
                    assert abs(cg.iterations - it) <= max(3, it // 50)
            # the captured graph replays the same kernels: same bits, same iteration count
            assert sols[(tr, True)][1] == sols[(tr, False)][1] and torch.equal(sols[(tr, True)][0], sols[(tr, False)][0]), tr
            # fixed-iteration mode (bench): recurrence residual == true residual b - A x
            cg = dist.DistCG(A, p, rel_tol=0.0, max_iter=25, op=op)
            x = torch.zeros_like(bl)
            cg.solve(bl, x, fixed_iters=25)
            r = op.vectors()[0][lo:hi].clone()
            ax = torch.zeros_like(bl)
            op.mult(x, ax)
            num = torch.tensor([((bl - ax)[lo:hi] - r).pow(2).sum().item(), bl[lo:hi].pow(2).sum().item()],
                               dtype=torch.float64, device="cuda")
            op.allreduce_sum(num)
            assert (num[0] / num[1]).sqrt().item() < 1e-12, (tr, num.tolist())
        # matrix-free operator over the ranks (each rank applies its local cells, owned rows complete, owned-dof dot)
        pa = fem.PAOperator(f, bcs=[fem.DirichletBC(p.bc, p.g)])
        cgp = dist.DistCG(A, p, rel_tol=1e-12, max_iter=4000, op=op, pa=pa)
        xp = torch.zeros_like(bl)
        cgp.solve(bl, xp)
        xsp = dist.gather_owned(p, xp)
        assert cgp.converged and abs(cgp.iterations - it) <= max(3, it // 50)
        if rank == 0:
            assert np.linalg.norm(xsp - want) / np.linalg.norm(want) < 1e-10
        if len(transports) == 2:   # both transports solve the same system: same iterates up to the all-reduce order
            a, c = sols[("p2p", True)], sols[("nccl", True)]
            assert abs(a[1] - c[1]) <= 1 and ((a[0] - c[0]).norm() / c[0].norm()).item() < 1e-10
        op.close()
        # distributed Newton (residual + lifting + tangent + PCG on strips, ghost update of u after every increment,
        # all-reduced norms) against the oracle's Newton loop on the global mesh, both convergence conventions
        dg, fg = fm.damage_band(mg), fm.body_force(mg)
        for conv_rule in ("mfem", "dolfinx"):
            wantu, it_o, norms_o = oracle.newton(mg.etype, mg.x, mg.xdofmap, mg.dofmap, Eg, 0.3, bcg, gg, dnod=dg,
                                                 fnod=fg.ravel(), convention=conv_rule, max_iter=20)
            ns = fem.NewtonSolver(fem.ElasticityForm(p.mesh, p.E, 0.3, d=fm.damage_band(p.mesh)), [fem.DirichletBC(p.bc, p.g)],
                                  f=fm.body_force(p.mesh), convention=conv_rule, max_iter=20, part=p)
            ul = ns.solve()
            us = dist.gather_owned(p, ul)
            assert ns.converged and ns.iterations == it_o, (conv_rule, ns.iterations, it_o)
            assert_norms_match(ns.residual_norms, norms_o)
            if rank == 0:
                assert np.linalg.norm(us - wantu) / np.linalg.norm(wantu) < 1e-9
            ns.close()
        out.put((rank, "ok " + "+".join(transports)))
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
        assert msg.startswith("ok"), f"rank {rank}: {msg}"
    print("transports exercised:", results[0][1])


def test_dist_world1_matches_oracle():
    """femb200_dist_pcg with a single rank (no transport): graph replay and eager launches give the same bits,
    the solution matches the oracle PCG, and the per-call row range leaves the plan's other users alone."""
    import torch
    from femb200 import dist, fem
    p = dist.strip_partition_device(NX, ROWS, 0, 1)
    f = fem.ElasticityForm(p.mesh, p.E)
    A = fem.create_matrix(f)
    fem.assemble_matrix(A, f, bcs=[fem.DirichletBC(p.bc, p.g)])
    mg = global_mesh(1)
    np.testing.assert_array_equal(p.mesh.x.cpu().numpy(), mg.x)
    Eg = fm.young_per_cell(mg.ncells)
    bcg, gg = fm.dirichlet_markers(mg)
    rowptr, colidx = oracle.build_pattern(mg.nnodes, mg.dofmap)
    vals = oracle.assemble_matrix(mg.etype, mg.x, mg.xdofmap, mg.dofmap, Eg, 0.3, rowptr, colidx, bc=bcg)
    b = np.where(bcg != 0, gg, 1.0)
    want, it, _, conv = oracle.pcg(rowptr, colidx, vals, b, rtol=1e-12, maxit=4000, jacobi=True)
    bl = fem.to_device(b, np.float64)
    xs = []
    for graph in (False, True):
        cg = dist.DistCG(A, p, rel_tol=1e-12, max_iter=4000, use_graph=graph)
        x = torch.zeros_like(bl)
        cg.solve(bl, x)
        assert cg.converged and abs(cg.iterations - it) <= max(3, it // 50)
        assert np.linalg.norm(x.cpu().numpy() - want) / np.linalg.norm(want) < 1e-10
        xs.append(x)
    assert torch.equal(xs[0], xs[1])
