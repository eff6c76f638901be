"""CPU tests that pin the oracle (oracle/fem_oracle.c): known answers derived from
the reference's formulas on the reference's only mesh, agreement of
independent code paths, and structural properties.  No GPU needed."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle
from femb200 import mesh as fm


def rel(a, b):
    return abs(a - b) / abs(b)


def test_young_table_and_lame(kat):
    E = oracle.E_table()
    sv = kat["survey_values"]
    assert E[1] == pytest.approx(sv["E_range_1"], rel=1e-15)
    assert E[2] == pytest.approx(sv["E_range_2"], rel=1e-15)
    np.testing.assert_allclose(fm.young_table(), E, rtol=0, atol=0)  # product-side table == oracle table
    for k, key in ((1, "lame_phys1"), (2, "lame_phys2")):
        lam, mu = oracle.lame(E[k], 0.3)
        assert lam == pytest.approx(sv[key][0], rel=1e-14)
        assert mu == pytest.approx(sv[key][1], rel=1e-14)


def test_square_known_answers(square, kat):
    x, tri, tag = square["x"], square["tri"], square["tag"]
    sv = kat["survey_values"]
    assert x.shape[0] == 62 and tri.shape[0] == 98
    E = oracle.E_table()[tag % 200]
    rowptr, colidx = oracle.build_pattern(62, tri)
    assert rowptr[-1] == sv["nnz"] == 1520
    vals = oracle.assemble_matrix(oracle.P1, x, tri, tri, E, 0.3, rowptr, colidx)
    K = sp.csr_matrix((vals, colidx, rowptr), shape=(124, 124))
    assert rel(np.sqrt((vals ** 2).sum()), sv["fro_norm"]) < 1e-13
    assert rel(K.diagonal().sum(), sv["trace"]) < 1e-13
    u = np.zeros(124)
    u[0::2], u[1::2] = 0.01 * x[:, 0], -0.003 * x[:, 1]
    assert rel(u @ (K @ u), sv["energy_closed_form"]) < 1e-12
    # rigid-body modes: translations and the infinitesimal rotation
    fro = np.sqrt((vals ** 2).sum())
    tx, ty, rot = np.zeros(124), np.zeros(124), np.zeros(124)
    tx[0::2] = 1
    ty[1::2] = 1
    rot[0::2], rot[1::2] = -x[:, 1], x[:, 0]
    for t in (tx, ty, rot):
        assert np.abs(K @ t).max() < 1e-15 * fro
    assert abs(K - K.T).max() < 1e-15 * fro
    # all triangles of square.msh are clockwise
    det = (x[tri[:, 1], 0] - x[tri[:, 0], 0]) * (x[tri[:, 2], 1] - x[tri[:, 0], 1]) \
        - (x[tri[:, 2], 0] - x[tri[:, 0], 0]) * (x[tri[:, 1], 1] - x[tri[:, 0], 1])
    assert np.all(det < 0)


def test_p1_three_code_paths_agree(square):
    """B.D.B^t (M.cc:699-704,886-887) == tensor-product blocks (M.cc:705-717,893-911)
    == generic quadrature loop == ufcx-signature shim."""
    x, tri = square["x"], square["tri"]
    rng = np.random.default_rng(1)
    for e in range(0, 98, 7):
        xv = x[tri[e]]
        lam, mu = oracle.lame(5e7 + 1e6 * e, 0.3)
        A1 = oracle.p1_grad_mfem(xv, lam, mu)
        A2 = oracle.p1_grad_mfem(xv, lam, mu, blocks=True)
        A3 = oracle.element_grad(oracle.P1, xv, lam, mu, layout=oracle.LAYOUT_COLMAJOR_BYNODES).reshape(6, 6).T
        s = np.abs(A1).max()
        assert np.abs(A1 - A2).max() < 2e-15 * s
        assert np.abs(A1 - A3).max() < 2e-15 * s
        # ufcx: w = [d(3), E(1), u(6)], c = [nu], coordinate_dofs 3x3
        Ee = 5e7 + 1e6 * e
        w = np.concatenate([np.zeros(3), [Ee], rng.standard_normal(6)])
        cd = np.zeros((3, 3))
        cd[:, :2] = xv
        A4 = oracle.tabulate_tensor_J_p1(w, np.array([0.3]), cd)
        # byNODES col-major -> interleaved row-major: A4[2a+i, 2b+k] = A1[i*3+a, k*3+b]
        perm = np.array([0, 3, 1, 4, 2, 5])
        assert np.abs(A4 - A1[np.ix_(perm, perm)]).max() < 2e-15 * s


def test_damaged_tangent_closed_form_vs_ad():
    """Closed form (M.cc:736-872) vs nested-dual Hessian of psi (M.cc:100-155,752-765):
    the reference documents agreement at 1e-15 (doc.tex:2215-2221)."""
    rng = np.random.default_rng(7)
    lam, mu = oracle.lame(7e7, 0.3)
    worst = 0.0
    for _ in range(200):
        g = rng.standard_normal((2, 2)) * 1e-3
        eps = 0.5 * (g + g.T)
        d = rng.uniform(0.05, 0.95)
        Dc = oracle.tangent(oracle.TANGENT_CLOSED, lam, mu, d, eps.ravel())
        Da = oracle.tangent(oracle.TANGENT_AD, lam, mu, d, eps.ravel())
        worst = max(worst, np.abs(Dc - Da).max() / np.abs(Dc).max())
    assert worst < 1e-12
    # d = 0 -> Hooke (M.cc:873-881)
    D0 = oracle.tangent(oracle.TANGENT_CLOSED, lam, mu, 0.0, np.zeros(4))
    np.testing.assert_allclose(D0, [[2 * mu + lam, lam, 0], [lam, 2 * mu + lam, 0], [0, 0, mu]], rtol=1e-15)
    # null strain with d > 0: (1 - d) Hooke (M.cc:861-870)
    Dn = oracle.tangent(oracle.TANGENT_CLOSED, lam, mu, 0.3, np.zeros(4))
    np.testing.assert_allclose(Dn, 0.7 * D0, rtol=1e-15)


@pytest.mark.parametrize("kind,n,expected_blocks", [
    ("P1", 6, None), ("P2", 5, 46 * 25 + 16 * 5 + 1), ("Q2", 5, 64 * 25 + 16 * 5 + 1)])
def test_pattern_counts_and_operator_properties(kind, n, expected_blocks):
    m = {"P1": lambda: fm.structured_triangles(n, order=1), "P2": lambda: fm.structured_triangles(n, order=2),
         "Q2": lambda: fm.structured_quads_q2(n)}[kind]()
    m = fm.jitter(m, 0.2, seed=3)
    et = m.etype
    rowptr, colidx = oracle.build_pattern(m.nnodes, m.dofmap)
    if expected_blocks is not None:
        assert rowptr[-1] == 4 * expected_blocks  # SURVEY.md 8d exact structural counts
    # columns ascending and unique in every row
    for r in range(0, 2 * m.nnodes, max(1, m.nnodes // 17)):
        c = colidx[rowptr[r]:rowptr[r + 1]]
        assert np.all(np.diff(c) > 0)
    E = fm.young_per_cell(m.ncells)
    vals = oracle.assemble_matrix(et, m.x, m.xdofmap, m.dofmap, E, 0.3, rowptr, colidx)
    K = sp.csr_matrix((vals, colidx, rowptr), shape=(m.ndofs, m.ndofs))
    fro = np.sqrt((vals ** 2).sum())
    assert abs(K - K.T).max() < 1e-14 * fro
    tx, ty, rot = np.zeros(m.ndofs), np.zeros(m.ndofs), np.zeros(m.ndofs)
    tx[0::2] = 1
    ty[1::2] = 1
    rot[0::2], rot[1::2] = -m.x[:, 1], m.x[:, 0]
    for t in (tx, ty, rot):
        assert np.abs(K @ t).max() < 1e-13 * fro
    # patch test: a linear field has zero internal force at interior nodes when E is uniform
    Eu = np.full(m.ncells, 3.0e7)
    vu = oracle.assemble_matrix(et, m.x, m.xdofmap, m.dofmap, Eu, 0.3, rowptr, colidx)
    Ku = sp.csr_matrix((vu, colidx, rowptr), shape=(m.ndofs, m.ndofs))
    u = np.zeros(m.ndofs)
    u[0::2] = 0.01 * m.x[:, 0] + 0.002 * m.x[:, 1]
    u[1::2] = -0.003 * m.x[:, 1] + 0.001 * m.x[:, 0]
    r = Ku @ u
    interior = (m.x[:, 0] > 1e-9) & (m.x[:, 0] < 1 - 1e-9) & (m.x[:, 1] > 1e-9) & (m.x[:, 1] < m.x[:, 1].max() - 1e-9)
    idx = np.nonzero(interior)[0]
    assert np.abs(np.concatenate([r[2 * idx], r[2 * idx + 1]])).max() < 1e-12 * np.abs(r).max()
    # assembled SpMV == matrix-free apply (with and without Dirichlet)
    rng = np.random.default_rng(5)
    v = rng.standard_normal(m.ndofs)
    y_mf = oracle.apply_matrix_free(et, m.x, m.xdofmap, m.dofmap, E, 0.3, v)
    assert np.abs(K @ v - y_mf).max() < 1e-13 * np.abs(y_mf).max()
    bc, _ = fm.dirichlet_markers(m)
    vb = oracle.assemble_matrix(et, m.x, m.xdofmap, m.dofmap, E, 0.3, rowptr, colidx, bc=bc)
    y_b = oracle.spmv(rowptr, colidx, vb, v)
    y_mfb = oracle.apply_matrix_free(et, m.x, m.xdofmap, m.dofmap, E, 0.3, v, bc=bc)
    assert np.abs(y_b - y_mfb).max() < 1e-13 * np.abs(y_b).max()


def test_dirichlet_semantics(square):
    """F.cc:847-862: rows and columns of constrained dofs zeroed, 1.0 on the diagonal."""
    x, tri, tag = square["x"], square["tri"], square["tag"]
    E = oracle.E_table()[tag % 200]
    rowptr, colidx = oracle.build_pattern(62, tri)
    bc = np.zeros(124, dtype=np.uint8)
    left = np.nonzero(np.abs(x[:, 0]) < 1e-10)[0]
    right = np.nonzero(np.abs(x[:, 0] - 1) < 1e-10)[0]
    for nodes in (left, right):
        bc[2 * nodes] = bc[2 * nodes + 1] = 1
    v0 = oracle.assemble_matrix(oracle.P1, x, tri, tri, E, 0.3, rowptr, colidx)
    v1 = oracle.assemble_matrix(oracle.P1, x, tri, tri, E, 0.3, rowptr, colidx, bc=bc)
    K0 = sp.csr_matrix((v0, colidx, rowptr), shape=(124, 124)).toarray()
    K1 = sp.csr_matrix((v1, colidx, rowptr), shape=(124, 124)).toarray()
    c = bc.astype(bool)
    expect = K0.copy()
    expect[c, :] = 0
    expect[:, c] = 0
    expect[c, c] = 1.0
    np.testing.assert_array_equal(K1, expect)


def test_pcg_mfem_semantics():
    m = fm.jitter(fm.structured_triangles(8, order=2), 0.2, seed=11)
    E = fm.young_per_cell(m.ncells)
    rowptr, colidx = oracle.build_pattern(m.nnodes, m.dofmap)
    bc, g = fm.dirichlet_markers(m)
    vals = oracle.assemble_matrix(m.etype, m.x, m.xdofmap, m.dofmap, E, 0.3, rowptr, colidx, bc=bc)
    full = oracle.assemble_matrix(m.etype, m.x, m.xdofmap, m.dofmap, E, 0.3, rowptr, colidx)
    b = -oracle.spmv(rowptr, colidx, full, g)
    b[bc != 0] = g[bc != 0]
    xs, it, fn, conv = oracle.pcg(rowptr, colidx, vals, b, rtol=1e-12, maxit=2000, jacobi=True)
    assert conv and 0 < it < 2000
    K = sp.csr_matrix((vals, colidx, rowptr), shape=(m.ndofs, m.ndofs))
    ref = sp.linalg.spsolve(K.tocsc(), b)
    assert np.linalg.norm(xs - ref) / np.linalg.norm(ref) < 1e-9
    np.testing.assert_allclose(xs[bc != 0], g[bc != 0], rtol=0, atol=1e-13)
    # maxit reached: not converged, iteration count = maxit
    _, it2, _, conv2 = oracle.pcg(rowptr, colidx, vals, b, rtol=1e-12, maxit=5, jacobi=True)
    assert (not conv2) and it2 == 5
    # zero right-hand side converges in zero iterations
    _, it3, _, conv3 = oracle.pcg(rowptr, colidx, vals, np.zeros_like(b))
    assert conv3 and it3 == 0


def test_oracle_newton_consistency():
    """The residual is the derivative-consistent partner of the tangent: Newton converges in one
    step on the linear problem and quadratically with damage; the imposed values are reached."""
    m = fm.jitter(fm.structured_triangles(6, order=2), 0.2, seed=3)
    E = fm.young_per_cell(m.ncells)
    bc, g = fm.dirichlet_markers(m)
    f = fm.body_force(m).ravel()
    u0, it0, n0 = oracle.newton(m.etype, m.x, m.xdofmap, m.dofmap, E, 0.3, bc, g, fnod=f)
    assert it0 == 1 and n0[1] < 1e-9 * n0[0]
    u1, it1, n1 = oracle.newton(m.etype, m.x, m.xdofmap, m.dofmap, E, 0.3, bc, g, dnod=fm.damage_band(m), fnod=f)
    assert 2 <= it1 <= 6 and n1[-1] <= max(1e-7 * n1[0], 5e-8)
    assert n1[2] / n1[1] < 0.1 * n1[1] / n1[0] * 10  # super-linear decrease
    np.testing.assert_allclose(u1[bc != 0], g[bc != 0], atol=1e-14)
    # finite-difference check of the tangent against the residual (unconstrained, d > 0)
    rng = np.random.default_rng(0)
    u = 1e-3 * rng.standard_normal(m.ndofs)
    d = fm.damage_band(m)
    rowptr, colidx = oracle.build_pattern(m.nnodes, m.dofmap)
    K = sp.csr_matrix((oracle.assemble_matrix(m.etype, m.x, m.xdofmap, m.dofmap, E, 0.3, rowptr, colidx, dnod=d, u=u),
                       colidx, rowptr), shape=(m.ndofs, m.ndofs))
    v = rng.standard_normal(m.ndofs)
    h = 1e-7
    Fp = oracle.assemble_vector(m.etype, m.x, m.xdofmap, m.dofmap, E, 0.3, u + h * 1e-3 * v, dnod=d)
    Fm = oracle.assemble_vector(m.etype, m.x, m.xdofmap, m.dofmap, E, 0.3, u - h * 1e-3 * v, dnod=d)
    fd = (Fp - Fm) / (2 * h * 1e-3)
    assert np.linalg.norm(fd - K @ v) / np.linalg.norm(K @ v) < 1e-5
