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


def strip_partition_device(nx: int, ny_total: int, rank: int, world: int, jitter_amp: float = 0.2, seed: int = 1234,
                           device=None) -> StripPartition:
    """strip_partition for large meshes: the same rank-local P2 strip, generated with torch on `device`
    (lattice slicing instead of per-cell fancy indexing; 67 M cells in about a second).  Every array is
    bit-identical to the numpy version (same IEEE operations: linspace as index * step, the per-lattice-row
    Philox jitter drawn on the host, edge nodes as 0.5 * (a + b) of their two vertices); the fields of
    part.mesh / E / bc / g are torch tensors.  Checked against strip_partition in tests/test_dist_cpu.py."""
    if ny_total % world:
        raise ValueError("strip_partition: ny_total must be a multiple of the number of ranks")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    rows = ny_total // world
    r0, r1 = rank * rows, (rank + 1) * rows
    r1g = min(r1 + 1, ny_total)
    ny = r1g - r0
    mx, my = 2 * nx + 1, 2 * ny + 1
    f64, i32, i64 = torch.float64, torch.int32, torch.int64
    # numpy.linspace(0, stop, num): arange(num) * (stop / (num - 1)), last entry = stop
    xs = torch.arange(mx, dtype=f64, device=dev) * (1.0 / (mx - 1))
    xs[-1] = 1.0
    tot = 2 * ny_total + 1
    ys = torch.arange(2 * r0, 2 * r1g + 1, dtype=f64, device=dev) * ((ny_total / nx) / (tot - 1))
    if 2 * r1g == tot - 1:
        ys[-1] = ny_total / nx
    x = torch.empty((my, mx, 2), dtype=f64, device=dev)
    x[:, :, 0] = xs[None, :]
    x[:, :, 1] = ys[:, None]
    if jitter_amp != 0.0:
        h = 1.0 / nx
        vrows = [lr for lr in range(0, my, 2) if 0 < 2 * r0 + lr < tot - 1]
        if vrows:
            J = np.empty((len(vrows), nx + 1, 2))
            for k, lr in enumerate(vrows):
                d = np.random.default_rng([seed, 2 * r0 + lr]).uniform(-jitter_amp * h, jitter_amp * h, size=(nx + 1, 2))
                d[0] = d[-1] = 0.0
                J[k] = d
            Jd = torch.from_numpy(J).to(dev)
            idx = torch.tensor(vrows, dtype=i64, device=dev)
            x[idx, 0::2] += Jd
            del Jd, J
    # edge nodes: midpoint of their two vertices (horizontal, vertical, right diagonal)
    x[0::2, 1::2] = 0.5 * (x[0::2, 0:-1:2] + x[0::2, 2::2])
    x[1::2, 0::2] = 0.5 * (x[0:-1:2, 0::2] + x[2::2, 0::2])
    x[1::2, 1::2] = 0.5 * (x[0:-1:2, 0:-1:2] + x[2::2, 2::2])
    x = x.view(my * mx, 2)
    cx = torch.arange(nx, dtype=i64, device=dev)[None, :]
    cy = torch.arange(ny, dtype=i64, device=dev)[:, None]
    v0 = ((2 * cx) + mx * (2 * cy)).reshape(-1)
    v1, v2, v3 = v0 + 2, v0 + 2 * mx, v0 + 2 * mx + 2
    dm = torch.empty((2 * nx * ny, 6), dtype=i32, device=dev)
    for k, (a, b, c) in enumerate(((v0, v1, v3), (v0, v2, v3))):
        dm[k::2, 0], dm[k::2, 1], dm[k::2, 2] = a.to(i32), b.to(i32), c.to(i32)
        dm[k::2, 3], dm[k::2, 4], dm[k::2, 5] = ((b + c) // 2).to(i32), ((a + c) // 2).to(i32), ((a + b) // 2).to(i32)
    del v0, v1, v2, v3
    tri = dm[:, :3].contiguous()
    cell_offset = 2 * nx * r0
    tab = torch.from_numpy(fm.young_table()).to(dev)
    E = tab[torch.arange(cell_offset, cell_offset + dm.shape[0], dtype=i64, device=dev) % 200]
    bc = torch.zeros((my, mx, 2), dtype=torch.uint8, device=dev)
    bc[:, 0, :] = 1
    bc[:, -1, :] = 1
    g = torch.zeros((my, mx, 2), dtype=f64, device=dev)
    g[:, -1, 0] = 0.01
    m = fm.Mesh(fm.P2, x, tri, dm, nx, ny, {"kind": "structured-tri-right", "order": 2, "jitter": jitter_amp})
    own_row_lo = 0 if rank == 0 else 1
    own_row_hi = 2 * rows + 1
    part = StripPartition(rank, world, nx, ny_total, m, E, bc.view(-1), g.view(-1), own_row_lo * mx, own_row_hi * mx,
                          2 * nx * rows, 2 * r0 * mx, cell_offset)
    if rank > 0:
        part.sends[rank - 1] = (1 * mx, 3 * mx)
        part.recvs[rank - 1] = (0, mx)
    if rank < world - 1:
        part.sends[rank + 1] = ((2 * rows) * mx, (2 * rows + 1) * mx)
        part.recvs[rank + 1] = ((2 * rows + 1) * mx, (2 * rows + 3) * mx)
    return part


def trivial_partition(mesh: fm.Mesh) -> StripPartition:
    """The whole mesh as the one partition of a single rank (no neighbours, every node owned)."""
    return StripPartition(0, 1, mesh.nx, mesh.ny, mesh, None, None, None, 0, mesh.nnodes, mesh.ncells, 0, 0)


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


class DistOperator:
    """The rank-local piece of the distributed operator: a femb200_dist communicator (C, csrc/dist.cu) over the
    plan of `A` with the partition's owned range and neighbour table, plus its transport.

    transport = "p2p":  NVLink peer memory (CUDA IPC): halo = one kernel storing into the neighbours' ghost
                        rows, dot-product all-reduce fused into the CG's scalar kernel -- no collective launches;
                "nccl": grouped ncclSend/ncclRecv + ncclAllReduce on a communicator the library creates from a
                        unique id broadcast through torch.distributed;
                "auto": p2p when every rank can map every peer, else nccl.
    torch.distributed (any backend) is only the bootstrap: it carries the 128-byte NCCL id and the 256-byte
    IPC blobs.  All ranks must construct, use and close the operator in the same order."""

    def __init__(self, A, part: StripPartition, transport: str = "auto", group=None):
        self.A, self.part, self.group = A, part, group
        self._d = C.c_void_p()
        self._comm = C.c_void_p()
        nb = sorted(part.sends)
        if sorted(part.recvs) != nb:
            raise ValueError("DistOperator: sends and receives must name the same neighbours")
        peers = np.array(nb, dtype=np.int32)
        sl = np.array([part.sends[p][0] for p in nb], dtype=np.int64)
        sh = np.array([part.sends[p][1] for p in nb], dtype=np.int64)
        rl = np.array([part.recvs[p][0] for p in nb], dtype=np.int64)
        rh = np.array([part.recvs[p][1] for p in nb], dtype=np.int64)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        capi.call("femb200_dist_create", A.plan, part.rank, part.world, part.own_lo, part.own_hi, len(nb), ptr(peers),
                  ptr(sl), ptr(sh), ptr(rl), ptr(rh), _st(), C.byref(self._d))
        self.transport = "none"
        if part.world > 1:
            if transport in ("auto", "p2p"):
                ok = self._attach_p2p()
                if not ok and transport == "p2p":
                    raise RuntimeError("DistOperator: peer memory could not be mapped on every rank")
            if self.transport == "none":
                self._attach_nccl()

    # -- bootstrap ---------------------------------------------------------------------------
    def _attach_p2p(self) -> bool:
        blob = (C.c_ubyte * capi.DIST_BLOB_BYTES)()
        capi.call("femb200_dist_p2p_export", self._d, blob)
        blobs = [None] * self.part.world
        td.all_gather_object(blobs, bytes(blob), group=self.group)
        buf = (C.c_ubyte * (capi.DIST_BLOB_BYTES * self.part.world)).from_buffer_copy(b"".join(blobs))
        rc = capi.lib().femb200_dist_p2p_attach(self._d, buf)
        self.p2p_error = None if rc == 0 else capi.lib().femb200_last_error().decode(errors="replace")
        oks = [None] * self.part.world
        td.all_gather_object(oks, rc == 0, group=self.group)
        if all(oks):
            self.transport = "p2p"
            return True
        return False

    def _attach_nccl(self):
        idb = (C.c_ubyte * 128)()
        if self.part.rank == 0:
            capi.call("femb200_dist_nccl_unique_id", idb)
        box = [bytes(idb)]
        td.broadcast_object_list(box, src=0, group=self.group)
        idb = (C.c_ubyte * 128).from_buffer_copy(box[0])
        capi.call("femb200_dist_nccl_comm_create", idb, self.part.rank, self.part.world, C.byref(self._comm))
        capi.call("femb200_dist_attach_nccl", self._d, self._comm)
        capi.call("femb200_dist_set_transport", self._d, capi.DIST_NCCL)
        self.transport = "nccl"

    def use(self, transport: str):
        """Switch between attached transports ("p2p" / "nccl"); attaches NCCL on first use."""
        if transport == "nccl" and not self._comm.value:
            self._attach_nccl()
            return
        capi.call("femb200_dist_set_transport", self._d, capi.DIST_P2P if transport == "p2p" else capi.DIST_NCCL)
        self.transport = transport

    def close(self):
        if getattr(self, "_d", None) is not None and self._d.value:
            torch.cuda.synchronize()
            if self.part.world > 1:
                td.barrier(group=self.group)          # nobody stores into an arena that is about to go
            capi.lib().femb200_dist_destroy(self._d)
            self._d = C.c_void_p()
            if self._comm.value:
                capi.lib().femb200_dist_nccl_comm_destroy(self._comm)
                self._comm = C.c_void_p()

    def __del__(self):
        try:
            if getattr(self, "_d", None) is not None and self._d.value and self.part.world == 1:
                capi.lib().femb200_dist_destroy(self._d)
                self._d = C.c_void_p()
        except Exception:
            pass

    @property
    def handle(self):
        return self._d

    # -- collectives / operator -----------------------------------------------------------------
    def allreduce_sum(self, v: torch.Tensor) -> torch.Tensor:
        """In-place sum over the ranks of a device vector of at most 3 doubles."""
        capi.call("femb200_dist_allreduce_sum", self._d, _p(v), v.numel(), _st())
        return v

    def halo(self, v: torch.Tensor) -> torch.Tensor:
        capi.call("femb200_dist_halo", self._d, _p(v), _st())
        return v

    def mult(self, v: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """y[owned] = (A v)[owned] with the ghost update of v."""
        capi.call("femb200_dist_mult", self._d, _p(self.A.values), _p(v), _p(y), _st())
        return y

    def vectors(self):
        """(r, d, z, scal) of the last solve as torch views of the communicator's device memory."""
        ps = [C.c_void_p() for _ in range(4)]
        capi.call("femb200_dist_vectors", self._d, *[C.byref(p) for p in ps])
        n = 2 * self.part.mesh.nnodes
        return tuple(_view(p.value, k) for p, k in zip(ps, (n, n, n, capi.SC_COUNT)))


def _st():
    return torch.cuda.current_stream().cuda_stream


def _view(ptr: int, n: int) -> torch.Tensor:
    """float64 device tensor over memory owned by the C library (no copy; valid while the owner lives)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
    return torch.as_tensor(h, device="cuda")


class DistCG:
    """(Jacobi-)PCG over the ranks with mfem::CGSolver semantics (M.cc:1502,1525-1528): femb200_dist_pcg, the
    single-GPU kernels on the owned rows + ghost update of the search direction + all-reduce of the three dot
    products, the whole loop in C (one CUDA graph per iteration).  world = 1 works without a process group."""

    def __init__(self, A, part: StripPartition, rel_tol=1e-12, abs_tol=0.0, max_iter=2000, check_every=25,
                 jacobi=True, group=None, transport: str = "auto", use_graph: bool = True, op: DistOperator | None = None,
                 pa=None):
        """`pa` (a fem.PAOperator built on the rank's local mesh, Dirichlet included) switches the operator of the
        solve from the assembled CSR of `A` to the matrix-free apply; `A` still provides the plan of the communicator."""
        self.A, self.part, self.pa = A, part, pa
        self.rel_tol, self.abs_tol, self.max_iter, self.check_every = rel_tol, abs_tol, max_iter, check_every
        self.op = DistOperator(A, part, transport, group) if op is None else op
        self.use_graph = use_graph
        self.dinv = None
        if jacobi:
            self.update_preconditioner()
        self.iterations, self.final_norm, self.converged = 0, 0.0, False

    def update_preconditioner(self):
        diag = self.A.diagonal() if self.pa is None else self.pa.diagonal()
        self.dinv = torch.empty_like(diag)
        capi.call("femb200_jacobi_setup", diag.numel(), _p(diag), _p(self.dinv), _st())

    def mult(self, v: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return self.op.mult(v, y)

    def solve(self, b: torch.Tensor, x: torch.Tensor, fixed_iters: int = 0) -> torch.Tensor:
        it, fn, cv = C.c_int(), C.c_double(), C.c_int()
        kind, opp, vals = (capi.OP_CSR, None, _p(self.A.values)) if self.pa is None else (capi.OP_PA, self.pa.handle, None)
        capi.call("femb200_dist_pcg", self.op.handle, kind, opp, vals, _p(b), _p(x), self.rel_tol,
                  self.abs_tol, self.max_iter, _p(self.dinv), self.check_every, int(fixed_iters), int(self.use_graph),
                  C.byref(it), C.byref(fn), C.byref(cv), _st())
        self.iterations, self.final_norm, self.converged = it.value, fn.value, bool(cv.value)
        return x

    def close(self):
        self.op.close()


def gather_owned(part: StripPartition, v_local: torch.Tensor, group=None) -> np.ndarray | None:
    """Owned dofs of every rank concatenated in global order on rank 0 (tests, output)."""
    mine = v_local[2 * part.own_lo:2 * part.own_hi].detach().cpu().contiguous()
    if part.world == 1:
        return mine.numpy()
    parts = [None] * part.world if part.rank == 0 else None
    td.gather_object(mine.numpy(), parts, dst=0, group=group)
    return np.concatenate(parts) if part.rank == 0 else None
