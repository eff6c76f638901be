"""include/femb200_mfem.hpp, the header-only MFEM adaptor of the C ABI (MFEM integrator role):
compiled against the MFEM stand-in of oracle/ref_shim (the real mfem.hpp is not in this image), linked with
libfemb200.so, then -- on the GPU -- run element by element next to the reference's OWN damIntegrator
(tests/golden/ref_p1_vectors.json, produced by compiling M.cc in place; oracle/_ref live when present)."""
import ctypes as C
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "fem-libraries_b200", "lib")
OUT = os.path.join(ROOT, "tests", "cxx", "_build")
SO = os.path.join(OUT, "libmfem_adaptor.so")


def build():
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(ROOT, "tests", "cxx", "mfem_adaptor_driver.cc")
    hdr = os.path.join(ROOT, "include", "femb200_mfem.hpp")
    if os.path.exists(SO) and os.path.getmtime(SO) > max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return SO
    cmd = [cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "oracle", "ref_shim"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"), src, "-o", SO,
           "-L", LIBDIR, "-lfemb200", "-L", os.path.join(cuda, "lib64"), "-lcudart", f"-Wl,-rpath,{LIBDIR}",
           f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}"]
    subprocess.check_call(cmd)
    return SO


def unhex(a):
    return np.array([float.fromhex(v) for v in a])


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_adaptor_compiles_and_links():
    """The header compiles as C++17 against the integrator base class and the dozen MFEM calls it uses, and the
    result links against the C ABI: every entry point of the driver resolves (no compute call without a GPU)."""
    L = C.CDLL(build())
    for name in ("adaptor_element_grad", "adaptor_element_vector", "adaptor_gradient_mult", "adaptor_last_error"):
        getattr(L, name)


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(ROOT, "tests", "golden", "ref_p1_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def lib():
    L = C.CDLL(build())
    dp = C.POINTER(C.c_double)
    L.adaptor_last_error.restype = C.c_char_p
    L.adaptor_element_grad.argtypes = [dp, C.c_double, C.c_double, C.c_double, dp, C.c_int, dp]
    L.adaptor_element_vector.argtypes = [dp, C.c_double, C.c_double, C.c_double, dp, dp, dp]
    return L


def P(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.gpu
def test_adaptor_element_grad_against_reference(gold, lib):
    """femb200::DamIntegrator::AssembleElementGrad (M.cc:639 signature, column-major byNODES elmat) against the
    reference's own integrator: d = 0 on triangles of square.msh, damaged cases incl. the special branches."""
    worst = 0.0
    for c in gold["linear"][::7]:
        xv = unhex(c["xv"])
        out = np.zeros(36)
        rc = lib.adaptor_element_grad(P(xv), float.fromhex(c["lam"]), float.fromhex(c["mu"]), 0.0, None, 0, P(out))
        assert rc == 0, lib.adaptor_last_error()
        worst = max(worst, rel(out, unhex(c["elmat_B"])))
    assert worst < 1e-13, worst
    for c in gold["damaged"]:
        xv, u = unhex(c["xv"]), unhex(c["elfun"])
        for variant in (0, 1):
            if variant == 1 and c["kind"] not in ("random", "null_strain"):
                continue
            out = np.zeros(36)
            rc = lib.adaptor_element_grad(P(xv), float.fromhex(c["lam"]), float.fromhex(c["mu"]), float.fromhex(c["d"]),
                                          P(u), variant, P(out))
            assert rc == 0, lib.adaptor_last_error()
            tol = 1e-6 if c["kind"] == "isotropic" else (1e-11 if variant else 1e-12)
            assert rel(out, unhex(c["elmat_B"])) < tol, (c["kind"], variant, rel(out, unhex(c["elmat_B"])))


@pytest.mark.gpu
def test_adaptor_element_vector_against_reference(gold, lib):
    """AssembleElementVector (M.cc:559 signature): stress term + load term with the coefficient given at the three
    points of the degree-2 rule, against the reference's elvect."""
    for c in gold["damaged"]:
        xv, u, fq = unhex(c["xv"]), unhex(c["elfun"]), unhex(c["fq"])
        out = np.zeros(6)
        rc = lib.adaptor_element_vector(P(xv), float.fromhex(c["lam"]), float.fromhex(c["mu"]), float.fromhex(c["d"]), P(u),
                                        P(fq), P(out))
        assert rc == 0, lib.adaptor_last_error()
        want = unhex(c["elvect"])
        scale = max(np.linalg.norm(want), 1e-6 * float.fromhex(c["lam"]) * 1e-3)
        tol = 1e-6 if c["kind"] == "isotropic" else 1e-12
        assert np.linalg.norm(out - want) / scale < tol, (c["kind"], out, want)


@pytest.mark.gpu
def test_adaptor_gradient_operator(lib):
    """femb200::GradientOperator (GetGradient + Mult on the device) against the oracle's assembled SpMV."""
    from oracle import oracle
    from femb200 import mesh as fm
    m = fm.jitter(fm.structured_triangles(9, 7, order=2), 0.2, seed=3)
    E = fm.young_per_cell(m.ncells)
    bc, _ = fm.dirichlet_markers(m)
    rowptr, colidx = oracle.build_pattern(m.nnodes, m.dofmap)
    vals = oracle.assemble_matrix(m.etype, m.x, m.xdofmap, m.dofmap, E, 0.3, rowptr, colidx, bc=bc)
    x = np.random.default_rng(1).standard_normal(m.ndofs)
    want = oracle.spmv(rowptr, colidx, vals, x)
    y = np.zeros(m.ndofs)
    nnz = C.c_long()
    ip = C.POINTER(C.c_int)
    lib.adaptor_gradient_mult.argtypes = [C.c_int, C.c_long, C.c_long, C.POINTER(C.c_double), ip, ip, C.POINTER(C.c_double),
                                          C.c_double, C.POINTER(C.c_ubyte), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.POINTER(C.c_long)]
    u = np.zeros(m.ndofs)
    xy = np.ascontiguousarray(m.x[:, :2])
    rc = lib.adaptor_gradient_mult(1, m.nnodes, m.ncells, P(xy), m.dofmap.ctypes.data_as(ip), m.xdofmap.ctypes.data_as(ip), P(E),
                                   0.3, bc.ctypes.data_as(C.POINTER(C.c_ubyte)), P(u), P(x), P(y), C.byref(nnz))
    assert rc == 0, lib.adaptor_last_error()
    assert nnz.value == rowptr[-1]
    assert rel(y, want) < 1e-12
