"""Multi-GPU path: element-block (strip) partition, halo exchange, distributed PCG.

Role in the reference: the MPI domain decomposition both libraries use (METIS /
ParMETIS partitions, ghost mode `none`, vertex ownership = lowest rank:
doc.tex:393-464, F.cc:159), the forward ghost update before an operator apply
(VecGhostUpdate(INSERT, FORWARD), F.cc:865-866; HypreParMatrix::Mult's halo) and
the MPI_Allreduce of the CG dot products inside CGSolver / KSP.

B200 design: one process per GPU (torch.distributed, NCCL over NVLink).  The
structured mesh is cut into horizontal strips of cell rows; rank r also integrates
the ONE row of neighbour cells that touches the interface nodes it owns (overlap
by one element), so assembly needs no communication at all and every owned matrix
row is complete.  Local node numbering stays lexicographic on the local lattice
(ghost row below, owned rows, two ghost rows above): owned nodes form one
contiguous range and every halo message is a contiguous slice.  Per CG iteration:
one grouped send/recv of the interface dofs of the search direction (P2, n = 1448:
2 x 2897 nodes x 16 B = 93 KB per neighbour) and two all-reduces of one double,
consumed on the device (no host synchronisation).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.distributed as td

from . import _capi as capi
from . import mesh as fm


def jitter_rows(mesh: fm.Mesh, amp: float, seed: int, first_lattice_row: int, total_lattice_rows: int) -> fm.Mesh:
    """Jitter of the interior VERTEX nodes of a P2/Q2 lattice, reproducible from the
    GLOBAL lattice row alone (one Philox stream per row), so that every rank
    generates exactly the coordinates the single-process mesh has."""
    if amp == 0.0:
        return mesh
    nx = mesh.nx
    mx = 2 * nx + 1
    h = 1.0 / nx
    x = mesh.x.copy()
    nrows = x.shape[0] // mx
    for lr in range(nrows):
        gr = first_lattice_row + lr
        if gr % 2 or gr == 0 or gr == total_lattice_rows - 1:
            continue
        rng = np.random.default_rng([seed, gr])
        d = rng.uniform(-amp * h, amp * h, size=(nx + 1, 2))
        d[0] = d[-1] = 0.0
        x[lr * mx:(lr + 1) * mx:2] += d
    dm = mesh.dofmap.astype(np.int64)
    if mesh.etype == fm.P2:
        for loc, (p, q) in zip((3, 4, 5), ((1, 2), (0, 2), (0, 1))):
            x[dm[:, loc]] = 0.5 * (x[dm[:, p]] + x[dm[:, q]])
    elif mesh.etype == fm.Q2:
        for loc, vs in ((1, (0, 2)), (3, (0, 6)), (5, (2, 8)), (7, (6, 8)), (4, (0, 2, 6, 8))):
            x[dm[:, loc]] = np.mean([x[dm[:, v]] for v in vs], axis=0)
    return fm.Mesh(mesh.etype, x, mesh.xdofmap, mesh.dofmap, mesh.nx, mesh.ny, dict(mesh.meta, jitter=amp))


@dataclass
class StripPartition:
    """Rank-local view of a structured P2 mesh cut into strips of cell rows."""
    rank: int
    world: int
    nx: int
    ny_total: int
    mesh: fm.Mesh          # owned cell rows + one ghost cell row above (if any)
    E: np.ndarray
    bc: np.ndarray
    g: np.ndarray
    own_lo: int            # owned local nodes = [own_lo, own_hi)
    own_hi: int
    n_owned_cells: int
    node_offset: int       # global node = local node + node_offset
    cell_offset: int       # global cell = local cell + cell_offset
    sends: dict = field(default_factory=dict)   # peer -> (lo, hi) local node range (or int32 node-id tensor) to send
    recvs: dict = field(default_factory=dict)   # peer -> (lo, hi) local node range to receive into

    @property
    def n_owned(self) -> int:
        return self.own_hi - self.own_lo

    @property
    def nnodes_global(self) -> int:
        return (2 * self.nx + 1) * (2 * self.ny_total + 1)

    def owned_nnz_blocks(self, A) -> int:
        brp, _ = A.block_csr()
        return int((brp[self.own_hi] - brp[self.own_lo]).item())


def strip_partition(nx: int, ny_total: int, order: int, rank: int, world: int, jitter_amp: float = 0.2,
                    seed: int = 1234) -> StripPartition:
    """Rank `rank` of `world` strips: cell rows [rank*ny_total/world, (rank+1)*ny_total/world)
    plus one ghost row of cells above.  P2 triangles (order 2) only."""
    if order != 2:
        raise ValueError("strip_partition: P2 triangles only")
    if ny_total % world:
        raise ValueError("strip_partition: ny_total must be a multiple of the number of ranks")
    rows = ny_total // world
    r0, r1 = rank * rows, (rank + 1) * rows
    r1g = min(r1 + 1, ny_total)                   # with the ghost cell row
    mx = 2 * nx + 1
    h = 1.0 / nx
    m = fm.structured_triangles(nx, r1g - r0, order=2, ly=(r1g - r0) * h, y0=r0 * h)
    # coordinates bit-identical to the single-process mesh: take the y of the GLOBAL lattice rows
    yg = np.linspace(0.0, ny_total / nx, 2 * ny_total + 1)[2 * r0:2 * r1g + 1]
    m.x[:, 1] = np.repeat(yg, mx)
    m = jitter_rows(m, jitter_amp, seed, 2 * r0, 2 * ny_total + 1)
    cell_offset = 2 * nx * r0
    E = fm.young_per_cell(m.ncells, first_cell=cell_offset)
    bc, g = fm.dirichlet_markers(m)
    first_row = 2 * r0
    own_row_lo = 0 if rank == 0 else 1            # the bottom lattice row belongs to the rank below
    own_row_hi = 2 * rows + 1                      # exclusive, local lattice rows
    part = StripPartition(rank, world, nx, ny_total, m, E, bc, g, own_row_lo * mx, own_row_hi * mx, 2 * nx * rows,
                          first_row * mx, cell_offset)
    if rank > 0:
        part.sends[rank - 1] = (1 * mx, 3 * mx)            # my first two owned rows = its two top ghost rows
        part.recvs[rank - 1] = (0, mx)                     # my bottom ghost row = its top owned row
    if rank < world - 1:
        part.sends[rank + 1] = ((2 * rows) * mx, (2 * rows + 1) * mx)
        part.recvs[rank + 1] = ((2 * rows + 1) * mx, (2 * rows + 3) * mx)
    return part


class Halo:
    """Forward ghost update of a local dof vector (2 dofs per node, blocked).  Works on
    CUDA tensors (NCCL) and on CPU tensors (gloo, used by the CPU tests of the host
    logic); messages are contiguous node ranges, packed with femb200_gather when a
    send list is not a range."""

    def __init__(self, part: StripPartition, group=None):
        self.part, self.group = part, group

    def forward(self, v: torch.Tensor) -> None:
        if self.part.world == 1:
            return
        ops, keep = [], []
        for peer, what in sorted(self.part.sends.items()):
            if isinstance(what, tuple):
                lo, hi = what
                buf = v[2 * lo:2 * hi]      # contiguous slice: no pack kernel needed
            else:                            # general send list (int32 node ids): pack
                idx = what.to(v.device)
                buf = torch.empty(2 * idx.numel(), dtype=v.dtype, device=v.device)
                if v.is_cuda:
                    capi.call("femb200_gather", idx.numel(), _p(idx), _p(v), _p(buf),
                              torch.cuda.current_stream().cuda_stream)
                else:
                    buf.view(-1, 2).copy_(v.view(-1, 2)[idx.long()])
            keep.append(buf)
            ops.append(td.P2POp(td.isend, buf, peer, self.group))
        for peer, (lo, hi) in sorted(self.part.recvs.items()):
            ops.append(td.P2POp(td.irecv, v[2 * lo:2 * hi], peer, self.group))
        for req in td.batch_isend_irecv(ops):
            req.wait()


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _off(t: torch.Tensor, n_doubles: int):
    return C.c_void_p(t.data_ptr() + 8 * n_doubles)


class DistCG:
    """(Jacobi-)PCG over the strips with mfem::CGSolver semantics (M.cc:1502,1525-1528):
    the single-GPU kernels of libfemb200 on the owned rows + halo exchange of the
    search direction + all-reduce of the three dot products."""

    def __init__(self, A, part: StripPartition, rel_tol=1e-12, abs_tol=0.0, max_iter=2000, check_every=25,
                 jacobi=True, group=None, overlap: bool = False):
        self.A, self.part = A, part
        self.rel_tol, self.abs_tol, self.max_iter, self.check_every = rel_tol, abs_tol, max_iter, check_every
        self.halo = Halo(part, group)
        self.group = group
        # Owned rows that read ghost values (strip layout: the two lattice rows above the bottom ghost row, the
        # lattice row below the top ghost rows) are applied after the halo exchange; all the other owned rows --
        # the plan's row range -- run while the exchange is in flight on a second stream.
        mx = 2 * part.nx + 1
        self.bottom = (part.own_lo, min(part.own_lo + 2 * mx, part.own_hi)) if part.rank > 0 else (part.own_lo, part.own_lo)
        self.top = (max(part.own_hi - mx, self.bottom[1]), part.own_hi) if part.rank < part.world - 1 else (part.own_hi, part.own_hi)
        # (measured at 2 GPUs: 0.854 ms per iteration with the overlap, 0.848 ms without -- the communication
        # cost of an iteration is the latency of its three all-reduces, not the 93 KB halo -- hence off by default)
        self.overlap = bool(overlap) and part.world > 1 and A.values.is_cuda
        if self.overlap:
            A.set_row_range(self.bottom[1], self.top[0])
            self.side = torch.cuda.Stream()
            self.ev_ready, self.ev_halo = torch.cuda.Event(), torch.cuda.Event()
        else:
            A.set_row_range(part.own_lo, part.own_hi)
        nl = 2 * part.mesh.nnodes
        dev = A.values.device
        self.r = torch.zeros(nl, dtype=torch.float64, device=dev)
        self.d = torch.zeros(nl, dtype=torch.float64, device=dev)
        self.z = torch.zeros(nl, dtype=torch.float64, device=dev)
        self.scal = torch.zeros(capi.SC_COUNT, dtype=torch.float64, device=dev)
        self.dinv = None
        if jacobi:
            diag = A.diagonal()
            self.dinv = torch.empty_like(diag)
            capi.call("femb200_jacobi_setup", diag.numel(), _p(diag), _p(self.dinv), self._st())
        self.iterations, self.final_norm, self.converged = 0, 0.0, False

    @staticmethod
    def _st():
        return torch.cuda.current_stream().cuda_stream

    def _allreduce(self, idx: int):
        if self.part.world > 1:
            td.all_reduce(self.scal[idx:idx + 1], group=self.group)

    def mult(self, v: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """y[owned] = (A v)[owned] with the ghost update of v: halo exchange on the second stream while the
        interior rows run, then the rows next to the ghosts (the operator apply of one CG iteration)."""
        A, st = self.A, self._st
        if not self.overlap:
            self.halo.forward(v)
            capi.call("femb200_spmv", A.plan, _p(A.values), _p(v), _p(y), st())
            return y
        main = torch.cuda.current_stream()
        self.ev_ready.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_ready)
            self.halo.forward(v)
            self.ev_halo.record(self.side)
        capi.call("femb200_spmv", A.plan, _p(A.values), _p(v), _p(y), st())
        main.wait_event(self.ev_halo)
        for lo, hi in (self.bottom, self.top):
            if hi > lo:
                capi.call("femb200_spmv_rows", A.plan, _p(A.values), _p(v), _p(y), lo, hi, None, 0, None, st())
        return y

    def solve(self, b: torch.Tensor, x: torch.Tensor, fixed_iters: int = 0) -> torch.Tensor:
        p = self.part
        o, n = 2 * p.own_lo, 2 * p.n_owned
        st = self._st
        A, scal = self.A, self.scal
        dinv_o = None if self.dinv is None else _off(self.dinv, o)
        capi.call("femb200_cg_set_tolerances", _p(scal), self.rel_tol, self.abs_tol, st())
        capi.call("femb200_cg_init", n, _off(b, o), dinv_o, _off(x, o), _off(self.r, o), _off(self.d, o), _p(scal), st())
        self._allreduce(capi.SC_RED_NOM)
        capi.call("femb200_cg_scalar_step", _p(scal), 0, st())

        def apply():
            if not self.overlap:
                self.halo.forward(self.d)
                capi.call("femb200_cg_apply", A.plan, capi.OP_CSR, None, _p(A.values), _p(self.d), _p(self.z), _p(scal), st())
            else:
                main = torch.cuda.current_stream()
                self.ev_ready.record(main)                       # the search direction is final
                with torch.cuda.stream(self.side):
                    self.side.wait_event(self.ev_ready)
                    self.halo.forward(self.d)                    # NCCL send/recv of the interface rows
                    self.ev_halo.record(self.side)
                # interior rows: Ad and the partial <d, A d>, while the halo is in flight
                capi.call("femb200_cg_apply", A.plan, capi.OP_CSR, None, _p(A.values), _p(self.d), _p(self.z), _p(scal), st())
                main.wait_event(self.ev_halo)
                flag = _off(scal, capi.SC_FLAG)
                for lo, hi in (self.bottom, self.top):           # rows that read ghost values; dot accumulated
                    if hi > lo:
                        capi.call("femb200_spmv_rows", A.plan, _p(A.values), _p(self.d), _p(self.z), lo, hi,
                                  _off(scal, capi.SC_RED_DEN), 1, flag, st())
            self._allreduce(capi.SC_RED_DEN)
            capi.call("femb200_cg_scalar_step", _p(scal), 1, st())

        apply()
        nit = fixed_iters if fixed_iters > 0 else self.max_iter
        stopped = False
        i = 0
        while i < nit and not stopped:
            i += 1
            capi.call("femb200_cg_update_xr", n, _p(scal), _off(self.d, o), _off(self.z, o), dinv_o, _off(x, o),
                      _off(self.r, o), st())
            self._allreduce(capi.SC_RED_BETA)
            capi.call("femb200_cg_scalar_step", _p(scal), 2, st())
            if i < nit:
                capi.call("femb200_cg_update_dir", n, _p(scal), _off(self.r, o), dinv_o, _off(self.d, o), st())
                apply()
            if fixed_iters <= 0 and i % self.check_every == 0 and i < nit:
                stopped = scal[capi.SC_FLAG].item() != 0.0
        hs = scal.cpu().numpy()
        self.converged = hs[capi.SC_FLAG] == 1.0
        self.iterations = int(hs[capi.SC_ITERS]) if (self.converged or fixed_iters > 0) else self.max_iter
        self.final_norm = float(np.sqrt(max(hs[capi.SC_FINAL], 0.0)))
        return x


def gather_owned(part: StripPartition, v_local: torch.Tensor, group=None) -> np.ndarray | None:
    """Owned dofs of every rank concatenated in global order on rank 0 (tests, output)."""
    mine = v_local[2 * part.own_lo:2 * part.own_hi].detach().cpu().contiguous()
    if part.world == 1:
        return mine.numpy()
    parts = [None] * part.world if part.rank == 0 else None
    td.gather_object(mine.numpy(), parts, dst=0, group=group)
    return np.concatenate(parts) if part.rank == 0 else None
