"""ctypes front-end of the CPU oracle (oracle/fem_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(fem-libraries_b200/femb200) never imports this module.

Parity status: the P1 element kernel is pinned against the reference's own compiled
code (oracle/_ref, tests/golden/ref_p1_vectors.json); P2/Q2, the CSR convention and
the Jacobi-PCG have no reference code or vectors ("parity unpinned" for those; see
the header of fem_oracle.c and DESIGN.md).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

P1, P2, Q2 = 0, 1, 2
LAYOUT_ROWMAJOR_INTERLEAVED, LAYOUT_COLMAJOR_BYNODES = 0, 1
TANGENT_CLOSED, TANGENT_AD = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "fem_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_build_pattern.restype = C.c_int64
        _LIB.orc_pcg.restype = C.c_int
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _l(a):
    return None if a is None else a.ctypes.data_as(_lp)


def _b(a):
    return None if a is None else a.ctypes.data_as(_bp)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def E_table() -> np.ndarray:
    out = np.empty(200)
    lib().orc_E_table(_d(out))
    return out


def lame(E: float, nu: float):
    l, m = C.c_double(), C.c_double()
    lib().orc_lame(C.c_double(E), C.c_double(nu), C.byref(l), C.byref(m))
    return l.value, m.value


def tangent(variant, lam, mu, d, eps) -> np.ndarray:
    D = np.empty(9)
    e = _f64(np.asarray(eps).reshape(4))
    lib().orc_tangent(C.c_int(variant), C.c_double(lam), C.c_double(mu), C.c_double(d), _d(e), _d(D))
    return D.reshape(3, 3)


def stress(lam, mu, d, w, eps) -> np.ndarray:
    s = np.empty(4)
    e = _f64(np.asarray(eps).reshape(4))
    lib().orc_stress(C.c_double(lam), C.c_double(mu), C.c_double(d), C.c_double(w), _d(e), _d(s))
    return s.reshape(2, 2)


def element_grad(etype, xv, lam, mu, dnod=None, u=None, variant=TANGENT_CLOSED,
                 layout=LAYOUT_ROWMAJOR_INTERLEAVED) -> np.ndarray:
    nd = (3, 6, 9)[etype]
    A = np.zeros((2 * nd, 2 * nd))
    xv, dnod, u = _f64(xv), _f64(dnod), _f64(u)
    lib().orc_element_grad(C.c_int(etype), _d(xv), C.c_double(lam), C.c_double(mu), _d(dnod), _d(u), C.c_int(variant),
                           C.c_int(layout), _d(A))
    return A


def p1_grad_mfem(xv, lam, mu, d=0.0, elfun=None, variant=TANGENT_CLOSED, blocks=False) -> np.ndarray:
    """6x6 column-major, byNODES (MFEM layout); returned as a numpy array whose
    [r, c] is elmat(r, c)."""
    A = np.zeros(36)
    xv, elfun = _f64(xv), _f64(elfun)
    fn = lib().orc_p1_grad_mfem_blocks if blocks else lib().orc_p1_grad_mfem_B
    fn(_d(xv), C.c_double(lam), C.c_double(mu), C.c_double(d), _d(elfun), C.c_int(variant), _d(A))
    return A.reshape(6, 6).T.copy()


def tabulate_tensor_J_p1(w, c, coordinate_dofs, A=None) -> np.ndarray:
    """ufcx-signature call (batch of one); A accumulates."""
    A = np.zeros(36) if A is None else A
    w, c, cd = _f64(w), _f64(c), _f64(coordinate_dofs)
    lib().orc_tabulate_tensor_J_p1(_d(A), _d(w), _d(c), _d(cd), None, None)
    return A.reshape(6, 6)


def p1_element_vector(xv, lam, mu, d, u, fnod=None) -> np.ndarray:
    r = np.zeros(6)
    xv, u, fnod = _f64(xv), _f64(u), _f64(fnod)
    lib().orc_p1_element_vector(_d(xv), C.c_double(lam), C.c_double(mu), C.c_double(d), _d(u), _d(fnod), _d(r))
    return r


def element_vector(etype, xv, lam, mu, u, dnod=None, fnod=None) -> np.ndarray:
    nd = (3, 6, 9)[etype]
    r = np.zeros(2 * nd)
    xv, u, dnod, fnod = _f64(xv), _f64(u), _f64(dnod), _f64(fnod)
    lib().orc_element_vector(C.c_int(etype), _d(xv), C.c_double(lam), C.c_double(mu), _d(dnod), _d(u), _d(fnod), _d(r))
    return r


def assemble_vector(etype, x, xdofmap, dofmap, E, nu, u, dnod=None, fnod=None) -> np.ndarray:
    """Unconstrained residual b = F(u) (F.cc:825 / M.cc:559-637)."""
    x, E, u, dnod, fnod = _f64(x), _f64(E), _f64(u), _f64(dnod), _f64(fnod)
    xd, dm = _i32(xdofmap), _i32(dofmap)
    b = np.empty(2 * x.shape[0])
    lib().orc_assemble_vector(C.c_int(etype), C.c_int64(dm.shape[0]), C.c_int64(x.shape[0]), _d(x), _i(xd), _i(dm),
                              _d(E), C.c_double(nu), _d(dnod), _d(u), _d(fnod), _d(b))
    return b


def newton(etype, x, xdofmap, dofmap, E, nu, bc, g, dnod=None, fnod=None, variant=TANGENT_CLOSED, rel_tol=1e-7,
           abs_tol=5e-8, max_iter=10, cg_rtol=1e-12, cg_maxit=4000, convention="mfem"):
    """Newton loop of the reference (tolerances M.cc:1531-1543; BC treatment F.cc:817-862):
    b = F(u); b -= (-1) J[:, bc] (g - u)_bc; b[bc] = -(g - u)_bc; J with BC rows/cols zeroed and unit diagonal;
    solve J du = b with Jacobi-PCG; u <- u - du.  Convergence, convention "mfem": |b| <= max(rel_tol |b0|, abs_tol)
    (M.cc:1535-1541); "dolfinx": |b| < abs_tol or |b| / r0 < rel_tol with r0 = |du_0|, the norm of the first
    increment, known after the first update only (dolfinx 0.8 NewtonSolver as driven by F.cc:869-891;
    doc.tex:2065-2068).  Returns u, iterations, residual norms."""
    x = _f64(x)
    nn = x.shape[0]
    rowptr, colidx = build_pattern(nn, dofmap)
    u = np.zeros(2 * nn)
    c = np.asarray(bc) != 0
    norms = []
    r0 = None
    for it in range(max_iter + 1):
        b = assemble_vector(etype, x, xdofmap, dofmap, E, nu, u, dnod, fnod)
        full = assemble_matrix(etype, x, xdofmap, dofmap, E, nu, rowptr, colidx, dnod=dnod, u=u, variant=variant)
        w = np.where(c, g - u, 0.0)
        b = b + spmv(rowptr, colidx, full, w)
        b[c] = -(g - u)[c]
        norms.append(float(np.linalg.norm(b)))
        if convention == "mfem":
            done = norms[-1] <= max(rel_tol * norms[0], abs_tol)
        else:
            done = norms[-1] < abs_tol or (r0 is not None and norms[-1] / r0 < rel_tol)
        if done or it == max_iter:
            break
        vals = assemble_matrix(etype, x, xdofmap, dofmap, E, nu, rowptr, colidx, dnod=dnod, u=u, variant=variant, bc=bc)
        du, _, _, conv = pcg(rowptr, colidx, vals, b, rtol=cg_rtol, maxit=cg_maxit, jacobi=True)
        if it == 0:
            r0 = float(np.linalg.norm(du))
        u = u - du
    return u, len(norms) - 1, norms


def build_pattern(nnodes: int, dofmap: np.ndarray):
    dm = _i32(dofmap)
    ncells, nd = dm.shape
    rowptr = np.zeros(2 * nnodes + 1, dtype=np.int64)
    nnz = lib().orc_build_pattern(C.c_int64(nnodes), C.c_int64(ncells), C.c_int(nd), _i(dm), _l(rowptr), None)
    colidx = np.empty(nnz, dtype=np.int32)
    lib().orc_build_pattern(C.c_int64(nnodes), C.c_int64(ncells), C.c_int(nd), _i(dm), _l(rowptr), _i(colidx))
    return rowptr, colidx


def assemble_matrix(etype, x, xdofmap, dofmap, E, nu, rowptr, colidx, dnod=None, u=None, variant=TANGENT_CLOSED,
                    bc=None, diag=1.0, nthreads=1, values=None) -> np.ndarray:
    x, E, dnod, u = _f64(x), _f64(E), _f64(dnod), _f64(u)
    xd, dm = _i32(xdofmap), _i32(dofmap)
    nnodes, ncells = x.shape[0], dm.shape[0]
    if values is None:
        values = np.empty(int(rowptr[-1]))
    bcm = None if bc is None else np.ascontiguousarray(bc, dtype=np.uint8)
    lib().orc_assemble_matrix(C.c_int(etype), C.c_int64(ncells), C.c_int64(nnodes), _d(x), _i(xd), _i(dm), _d(E),
                              C.c_double(nu), _d(dnod), _d(u), C.c_int(variant), _b(bcm), C.c_double(diag),
                              _l(rowptr), _i(colidx), _d(values), C.c_int(nthreads))
    return values


def tabulate_batch(etype, x, xdofmap, dofmap, E, nu, dnod=None, u=None, variant=TANGENT_CLOSED,
                   layout=LAYOUT_ROWMAJOR_INTERLEAVED) -> np.ndarray:
    x, E, dnod, u = _f64(x), _f64(E), _f64(dnod), _f64(u)
    xd, dm = _i32(xdofmap), _i32(dofmap)
    ncells, nd = dm.shape
    out = np.empty((ncells, 2 * nd, 2 * nd))
    lib().orc_tabulate_batch(C.c_int(etype), C.c_int64(ncells), _d(x), _i(xd), _i(dm), _d(E), C.c_double(nu),
                             _d(dnod), _d(u), C.c_int(variant), C.c_int(layout), _d(out))
    return out


def spmv(rowptr, colidx, values, x, nthreads=1, y=None) -> np.ndarray:
    x = _f64(x)
    n = rowptr.shape[0] - 1
    y = np.empty(n) if y is None else y
    lib().orc_spmv(C.c_int64(n), _l(rowptr), _i(colidx), _d(values), _d(x), _d(y), C.c_int(nthreads))
    return y


def apply_matrix_free(etype, x, xdofmap, dofmap, E, nu, xin, bc=None, diag=1.0) -> np.ndarray:
    x, E, xin = _f64(x), _f64(E), _f64(xin)
    xd, dm = _i32(xdofmap), _i32(dofmap)
    y = np.empty(2 * x.shape[0])
    bcm = None if bc is None else np.ascontiguousarray(bc, dtype=np.uint8)
    lib().orc_apply_matrix_free(C.c_int(etype), C.c_int64(dm.shape[0]), C.c_int64(x.shape[0]), _d(x), _i(xd), _i(dm),
                                _d(E), C.c_double(nu), _b(bcm), C.c_double(diag), _d(xin), _d(y))
    return y


def pcg(rowptr, colidx, values, b, rtol=1e-12, atol=0.0, maxit=2000, jacobi=True, nthreads=1):
    """mfem::CGSolver semantics (M.cc:1502,1525-1528). Returns x, iters, final_norm, converged."""
    b = _f64(b)
    n = b.shape[0]
    x = np.zeros(n)
    it, fn = C.c_int(), C.c_double()
    conv = lib().orc_pcg(C.c_int64(n), _l(rowptr), _i(colidx), _d(values), _d(b), _d(x), C.c_double(rtol),
                         C.c_double(atol), C.c_int(maxit), C.c_int(1 if jacobi else 0), C.byref(it), C.byref(fn),
                         C.c_int(nthreads))
    return x, it.value, fn.value, bool(conv)


def vertex_graph(nverts: int, tri: np.ndarray):
    """Edge graph of a triangulation as CSR (adjptr int64, adj int32): the neighbours of every
    vertex, ascending, no self (role of the vertex_edge / edge_vertex tables, M.cc:1209-1213)."""
    t = np.asarray(tri, dtype=np.int64)
    a = np.concatenate([t[:, 0], t[:, 1], t[:, 2], t[:, 1], t[:, 2], t[:, 0]])
    b = np.concatenate([t[:, 1], t[:, 2], t[:, 0], t[:, 0], t[:, 1], t[:, 2]])
    key = np.unique(a * nverts + b)
    rows, cols = key // nverts, key % nverts
    adjptr = np.zeros(nverts + 1, dtype=np.int64)
    np.add.at(adjptr, rows + 1, 1)
    return np.cumsum(adjptr), cols.astype(np.int32)


def smooth_damage(nverts: int, tri: np.ndarray, d0: np.ndarray, niter: int = 8, thr: float = 0.01) -> np.ndarray:
    """Damage-field smoothing (M.cc:1258-1315, F.py:160-199) on the vertex graph of `tri`."""
    adjptr, adj = vertex_graph(nverts, tri)
    d = np.array(d0, dtype=np.float64, copy=True)
    lib().orc_smooth_damage(C.c_int64(nverts), _l(adjptr), _i(adj), _d(d), C.c_int(niter), C.c_double(thr))
    return d


def cell_strain_stress(etype, x, xdofmap, dofmap, E, nu, u, dnod=None):
    """DG0 strain and stress (xx, xy, yy) per cell at the centroid (M.cc:333-430,1551-1563)."""
    x, E, u, dnod = _f64(np.asarray(x)[:, :2]), _f64(E), _f64(u), _f64(dnod)
    xd, dm = _i32(xdofmap), _i32(dofmap)
    nc = dm.shape[0]
    strain, stress_ = np.empty((nc, 3)), np.empty((nc, 3))
    lib().orc_cell_strain_stress(C.c_int(etype), C.c_int64(nc), _d(x), _i(xd), _i(dm), _d(E), C.c_double(nu), _d(dnod),
                                 _d(u), _d(strain), _d(stress_))
    return strain, stress_
