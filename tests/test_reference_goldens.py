"""The oracle (and, on the GPU, the CUDA element kernel) against golden vectors
produced by RUNNING THE REFERENCE'S OWN CODE: damIntegrator / asym_stress of
MFEM/mechanic2d/asym_elasto_damage_model.cc compiled in place against a minimal
MFEM stand-in (oracle/ref_shim, tests/golden/make_ref_vectors.py).  This is what
pins the oracle for the P1 element kernel; P2 / Q2 have no reference code."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-12  # relative Frobenius norm (north_star)


def unhex(a):
    return np.array([float.fromhex(v) for v in a])


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(ROOT, "tests", "golden", "ref_p1_vectors.json")) as f:
        return json.load(f)


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_linear_tangent_against_reference(gold):
    """d = 0 on every triangle of square.msh: oracle B.D.B^t, oracle blocks, oracle
    generic loop == reference USE_B path == reference blocks path."""
    assert len(gold["linear"]) == 98
    worst = 0.0
    for c in gold["linear"]:
        xv = unhex(c["xv"]).reshape(3, 2)
        lam, mu = float.fromhex(c["lam"]), float.fromhex(c["mu"])
        kB, kK = unhex(c["elmat_B"]), unhex(c["elmat_blocks"])
        assert rel(kB, kK) < 1e-15
        for blocks in (False, True):
            got = oracle.p1_grad_mfem(xv, lam, mu, blocks=blocks).T.ravel()  # [r, c] -> column-major
            worst = max(worst, rel(got, kB))
        gen = oracle.element_grad(oracle.P1, xv, lam, mu, layout=oracle.LAYOUT_COLMAJOR_BYNODES).ravel()
        worst = max(worst, rel(gen, kB))
    assert worst < 1e-14, worst


def test_abs_det_convention_on_clockwise_triangles(gold, square):
    """square.msh lists its triangles clockwise; MFEM swaps vertices 0 and 1 at load, the
    oracle keeps the file order and uses |det J|: same matrix up to the
    permutation of the two swapped vertices."""
    x, tri = square["x"], square["tri"]
    for c in gold["linear"][::5]:
        e = c["cell"]
        assert c["vertices"] == [int(tri[e][1]), int(tri[e][0]), int(tri[e][2])]
        lam, mu = float.fromhex(c["lam"]), float.fromhex(c["mu"])
        k_file = oracle.p1_grad_mfem(x[tri[e]], lam, mu)          # file (clockwise) order
        perm = np.array([1, 0, 2, 4, 3, 5])                        # byNODES dofs of the swapped vertices
        k_ref = unhex(c["elmat_B"]).reshape(6, 6).T
        assert rel(k_file[np.ix_(perm, perm)], k_ref) < 1e-14


def test_damaged_tangent_and_residual_against_reference(gold):
    """d > 0: closed-form tangent (M.cc:736-872) and residual with asym_stress + load
    (M.cc:207-329, 559-637), including the special branches."""
    kinds = set()
    for c in gold["damaged"]:
        kinds.add(c["kind"])
        xv = unhex(c["xv"]).reshape(3, 2)
        lam, mu, d = float.fromhex(c["lam"]), float.fromhex(c["mu"]), float.fromhex(c["d"])
        u = unhex(c["elfun"])                                      # byNODES
        kB, kK = unhex(c["elmat_B"]), unhex(c["elmat_blocks"])
        assert rel(kK, kB) < 1e-13
        for blocks in (False, True):
            got = oracle.p1_grad_mfem(xv, lam, mu, d=d, elfun=u, blocks=blocks).T.ravel()
            assert rel(got, kB) < TOL, (c["kind"], rel(got, kB))
        # AD tangent (M.cc:752-765) == closed form, as the reference documents (doc.tex:2215-2221)
        if c["kind"] in ("random", "null_strain"):
            ad = oracle.p1_grad_mfem(xv, lam, mu, d=d, elfun=u, variant=oracle.TANGENT_AD).T.ravel()
            assert rel(ad, kB) < 1e-11, (c["kind"], rel(ad, kB))
        # residual: oracle takes interleaved dofs and nodal loads; the golden holds the load at
        # the three quadrature points, so compare the stress part and the load part separately
        ui = np.stack([u[:3], u[3:]], axis=1).ravel()
        r0 = oracle.p1_element_vector(xv, lam, mu, d, ui)
        want0 = unhex(c["elvect_noload"])
        want0_i = np.stack([want0[:3], want0[3:]], axis=1).ravel()
        scale = max(np.linalg.norm(want0_i), 1e-6 * lam * 1e-3)
        assert np.linalg.norm(r0 - want0_i) / scale < TOL, (c["kind"], r0, want0_i)
    assert kinds == {"random", "null_strain", "full_damage_traction", "shear_free", "isotropic"}


def test_load_term_against_reference(gold):
    """-sum_q w_q N f(q) with the degree-2 rule (M.cc:613-632): a nodal load interpolated to
    the reference's quadrature points must give the reference's load vector."""
    lp = unhex(gold["load_points"]).reshape(3, 2)
    c = gold["damaged"][0]
    xv = unhex(c["xv"]).reshape(3, 2)
    lam, mu, d = float.fromhex(c["lam"]), float.fromhex(c["mu"]), float.fromhex(c["d"])
    u = unhex(c["elfun"])
    ui = np.stack([u[:3], u[3:]], axis=1).ravel()
    load_ref = unhex(c["elvect"]) - unhex(c["elvect_noload"])      # byNODES, from the given f at the points
    fq = unhex(c["fq"]).reshape(3, 2)
    # the unique P1 nodal field that takes the values fq at the three points
    N = np.stack([1 - lp[:, 0] - lp[:, 1], lp[:, 0], lp[:, 1]], axis=1)
    fnod = np.linalg.solve(N, fq)
    got = oracle.p1_element_vector(xv, lam, mu, d, ui, fnod.ravel()) - oracle.p1_element_vector(xv, lam, mu, d, ui)
    got_n = np.concatenate([got[0::2], got[1::2]])
    assert rel(got_n, load_ref) < 1e-12


def test_live_reference_library_if_present(gold):
    """In the build container oracle/_ref exists: the goldens must be reproducible bit for bit."""
    so = os.path.join(ROOT, "oracle", "_ref", "libref_B.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (no reference tree on this box)")
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    L.ref_element_grad.argtypes = [dp, C.c_double, C.c_double, C.c_double, dp, dp]
    for c in gold["damaged"][:10]:
        xv, u = unhex(c["xv"]), unhex(c["elfun"])
        out = np.zeros(36)
        L.ref_element_grad(xv.ctypes.data_as(dp), float.fromhex(c["lam"]), float.fromhex(c["mu"]),
                           float.fromhex(c["d"]), u.ctypes.data_as(dp), out.ctypes.data_as(dp))
        np.testing.assert_array_equal(out, unhex(c["elmat_B"]))


@pytest.mark.gpu
def test_cuda_element_kernel_against_reference(gold):
    """The batched CUDA element kernel (mfem layout: column-major, byNODES) against the
    reference's own element matrices, linear and damaged."""
    from femb200 import fem, mesh as fm
    # linear: all 98 cells in one launch (E recovered from mu: E = 2 mu (1 + nu))
    xs = np.concatenate([unhex(c["xv"]).reshape(3, 2) for c in gold["linear"]])
    cells = np.arange(3 * 98, dtype=np.int32).reshape(98, 3)
    E = np.array([2 * float.fromhex(c["mu"]) * 1.3 for c in gold["linear"]])
    m = fm.Mesh(fm.P1, xs, cells, cells)
    got = fem.element_grad_batched(fem.ElasticityForm(m, E, 0.3)).cpu().numpy().reshape(98, 36)
    want = np.stack([unhex(c["elmat_B"]) for c in gold["linear"]])
    assert rel(got, want) < TOL
    assert max(rel(got[e], want[e]) for e in range(98)) < TOL
    # damaged: one cell per case, nodal damage = d at all three vertices, u interleaved
    nd = len(gold["damaged"])
    xs = np.concatenate([unhex(c["xv"]).reshape(3, 2) for c in gold["damaged"]])
    cells = np.arange(3 * nd, dtype=np.int32).reshape(nd, 3)
    E = np.array([2 * float.fromhex(c["mu"]) * 1.3 for c in gold["damaged"]])
    dn = np.repeat([float.fromhex(c["d"]) for c in gold["damaged"]], 3)
    u = np.concatenate([np.stack([unhex(c["elfun"])[:3], unhex(c["elfun"])[3:]], axis=1).ravel() for c in gold["damaged"]])
    m = fm.Mesh(fm.P1, xs, cells, cells)
    want = np.stack([unhex(c["elmat_B"]) for c in gold["damaged"]])
    for variant in (0, 1):
        got = fem.element_grad_batched(fem.ElasticityForm(m, E, 0.3, d=dn, u=u, variant=variant)).cpu().numpy()
        got = got.reshape(nd, 36)
        for k, c in enumerate(gold["damaged"]):
            if variant == 1 and c["kind"] not in ("random", "null_strain"):
                continue  # the AD Hessian is singular on the measure-zero special branches
            tol = TOL if variant == 0 else 1e-11
            assert rel(got[k], want[k]) < tol, (variant, c["kind"], rel(got[k], want[k]))
