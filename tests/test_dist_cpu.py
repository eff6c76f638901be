"""Host-side logic of the multi-GPU path on CPU: strip partition, owned/ghost
numbering, halo exchange and all-reduce plumbing with world_size 2 and 3 over gloo.
The local operators are applied with the oracle (the checker): what is under test
is the partition + exchange, whose result must equal the global operator."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from oracle import oracle
from femb200 import dist, mesh as fm

NX, ROWS = 6, 4


def global_problem(world):
    m = fm.structured_triangles(NX, ROWS * world, order=2, ly=ROWS * world / NX)
    m = dist.jitter_rows(m, 0.2, 1234, 0, 2 * ROWS * world + 1)
    E = fm.young_per_cell(m.ncells)
    return m, E


def test_partition_tiles_the_global_mesh():
    for world in (1, 2, 3):
        mg, Eg = global_problem(world)
        owned = np.zeros(mg.nnodes, dtype=int)
        cells = np.zeros(mg.ncells, dtype=int)
        for r in range(world):
            p = dist.strip_partition(NX, ROWS * world, 2, r, world)
            lm = p.mesh
            # local -> global maps are pure offsets (lexicographic numbering is preserved)
            np.testing.assert_array_equal(lm.x, mg.x[p.node_offset:p.node_offset + lm.nnodes])
            np.testing.assert_array_equal(lm.dofmap + p.node_offset, mg.dofmap[p.cell_offset:p.cell_offset + lm.ncells])
            np.testing.assert_array_equal(p.E, Eg[p.cell_offset:p.cell_offset + lm.ncells])
            owned[p.node_offset + p.own_lo:p.node_offset + p.own_hi] += 1
            cells[p.cell_offset:p.cell_offset + p.n_owned_cells] += 1
            # every cell touching an owned node is present locally (rows are complete)
            touch = np.isin(mg.dofmap, np.arange(p.node_offset + p.own_lo, p.node_offset + p.own_hi)).any(axis=1)
            assert touch[:p.cell_offset].sum() == 0 and touch[p.cell_offset + lm.ncells:].sum() == 0
            # send / receive ranges pair up
            for peer, (lo, hi) in p.sends.items():
                q = dist.strip_partition(NX, ROWS * world, 2, peer, world)
                rlo, rhi = q.recvs[r]
                assert hi - lo == rhi - rlo and lo + p.node_offset == rlo + q.node_offset
                assert p.own_lo <= lo and hi <= p.own_hi
        assert (owned == 1).all() and (cells == 1).all()  # each node / cell owned exactly once


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mg, Eg = global_problem(world)
        rowptr, colidx = oracle.build_pattern(mg.nnodes, mg.dofmap)
        vals = oracle.assemble_matrix(mg.etype, mg.x, mg.xdofmap, mg.dofmap, Eg, 0.3, rowptr, colidx)
        rng = np.random.default_rng(3)
        vg = rng.standard_normal(mg.ndofs)
        want = oracle.spmv(rowptr, colidx, vals, vg)

        p = dist.strip_partition(NX, ROWS * world, 2, rank, world)
        lm = p.mesh
        lrp, lci = oracle.build_pattern(lm.nnodes, lm.dofmap)
        lv = oracle.assemble_matrix(lm.etype, lm.x, lm.xdofmap, lm.dofmap, p.E, 0.3, lrp, lci)
        # global CSR identity on the owned rows: structure bit-exact, values identical
        r_lo, r_hi = 2 * p.own_lo, 2 * p.own_hi
        g_lo = 2 * (p.node_offset + p.own_lo)
        np.testing.assert_array_equal(np.diff(lrp[r_lo:r_hi + 1]), np.diff(rowptr[g_lo:g_lo + (r_hi - r_lo) + 1]))
        seg = slice(lrp[r_lo], lrp[r_hi])
        gseg = slice(rowptr[g_lo], rowptr[g_lo + (r_hi - r_lo)])
        np.testing.assert_array_equal(lci[seg] + 2 * p.node_offset, colidx[gseg])
        np.testing.assert_array_equal(lv[seg], vals[gseg])

        # halo: owned entries known, ghosts poisoned
        v = torch.full((2 * lm.nnodes,), float("nan"), dtype=torch.float64)
        v[r_lo:r_hi] = torch.from_numpy(vg[g_lo:g_lo + (r_hi - r_lo)])
        dist.Halo(p).forward(v)
        np.testing.assert_array_equal(v.numpy(), vg[2 * p.node_offset:2 * (p.node_offset + lm.nnodes)])
        y = oracle.spmv(lrp, lci, lv, v.numpy())
        np.testing.assert_array_equal(y[r_lo:r_hi], want[g_lo:g_lo + (r_hi - r_lo)])
        # dot product of the owned parts, all-reduced == global dot
        part = torch.tensor([float(v.numpy()[r_lo:r_hi] @ y[r_lo:r_hi])], dtype=torch.float64)
        td.all_reduce(part)
        assert abs(part.item() - vg @ want) <= 1e-12 * np.abs(vg * want).sum()
        full = dist.gather_owned(p, torch.from_numpy(y))
        if rank == 0:
            np.testing.assert_array_equal(full, want)
        # general (index-list) send path packs the same bytes as the range path
        if world > 1:
            p2 = dist.strip_partition(NX, ROWS * world, 2, rank, world)
            p2.sends = {peer: torch.arange(lo, hi, dtype=torch.int32) for peer, (lo, hi) in p2.sends.items()}
            v2 = torch.full((2 * lm.nnodes,), float("nan"), dtype=torch.float64)
            v2[r_lo:r_hi] = torch.from_numpy(vg[g_lo:g_lo + (r_hi - r_lo)])
            dist.Halo(p2).forward(v2)
            np.testing.assert_array_equal(v2.numpy(), v.numpy())
        out.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        import traceback
        out.put((rank, traceback.format_exc()))
        raise
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_and_distributed_spmv_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [out.get(timeout=180) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"
