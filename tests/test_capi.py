"""The C-ABI library loads and exports every symbol include/femb200.h declares; the
host layer refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "femb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(femb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from femb200 import capi
    assert os.path.exists(capi.LIB_PATH), "libfemb200.so is not built (run __graft_entry__.build())"
    L = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/femb200.h but not exported"
    # the ctypes table covers the header and nothing else
    assert sorted(capi.ALL_SYMBOLS) == names
    assert capi.lib().femb200_version() == 100


def header_prototypes():
    """name -> (return kind, [argument kinds]) parsed from include/femb200.h; kinds: i32, i64, f64, ptr, void."""
    src = open(os.path.join(ROOT, "include", "femb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)

    def kind(t):
        t = t.strip()
        if "*" in t:
            return "ptr"
        base = re.sub(r"\b(const|unsigned)\b", "", t).split()
        base = base[0] if base else ""
        return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "double": "f64", "void": "void"}[base]

    out = {}
    for ret, name, args in re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(femb200_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        argl = [a for a in (x.strip() for x in args.split(",")) if a and a != "void"]
        # drop the parameter name (last identifier) unless the declaration is a bare type
        kinds = []
        for a in argl:
            m = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)?$", a)
            kinds.append(kind(m.group(1) if m.group(1).strip() else a))
        out[name] = (kind(ret), kinds)
    return out


def ctypes_kind(t):
    if t is None:
        return "void"
    if t in (ctypes.c_int, ctypes.c_int32):
        return "i32"
    if t is ctypes.c_int64:
        return "i64"
    if t is ctypes.c_double:
        return "f64"
    if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "_type_") and issubclass(t, ctypes._Pointer):
        return "ptr"
    raise AssertionError(f"unmapped ctypes type {t}")


def test_ctypes_table_matches_header_prototypes():
    """The binding table of femb200/_capi.py against the PARSED prototypes of include/femb200.h: same functions,
    same arity, and every argument / return value of the same kind (32-bit int, 64-bit int, double, pointer)."""
    from femb200 import capi
    protos = header_prototypes()
    assert sorted(protos) == sorted(capi.ALL_SYMBOLS)
    for name, args in capi.SIGNATURES.items():
        ret, kinds = protos[name]
        assert ret == "i32", f"{name}: returns {ret} in the header, the table assumes an int status"
        assert [ctypes_kind(a) for a in args] == kinds, f"{name}: ctypes {[ctypes_kind(a) for a in args]} vs header {kinds}"
    for name, (args, res) in capi._SPECIAL.items():
        ret, kinds = protos[name]
        assert ctypes_kind(res) == ret, f"{name}: return {ctypes_kind(res)} vs header {ret}"
        assert [ctypes_kind(a) for a in args] == kinds, f"{name}: ctypes {[ctypes_kind(a) for a in args]} vs header {kinds}"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    from femb200 import fem, mesh
    m = mesh.structured_triangles(2, order=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fem.ElasticityForm(m, mesh.young_per_cell(m.ncells))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fem.tabulate_tensor(np.zeros((6, 6)), np.zeros(10), np.array([0.3]), np.zeros((3, 3)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fem-libraries_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert "oracle" not in s.replace("no oracle", ""), f"{f} mentions the oracle"


def test_mesh_generators():
    from femb200 import mesh
    m = mesh.structured_triangles(4, 3, order=2)
    assert m.nnodes == 9 * 7 and m.ncells == 24 and m.dofmap.shape == (24, 6)
    # edge nodes are the midpoints of their vertices (basix order: edge i opposite vertex i)
    for loc, (p, q) in zip((3, 4, 5), ((1, 2), (0, 2), (0, 1))):
        np.testing.assert_allclose(m.x[m.dofmap[:, loc]], 0.5 * (m.x[m.dofmap[:, p]] + m.x[m.dofmap[:, q]]), atol=1e-15)
    q = mesh.structured_quads_q2(3)
    assert q.nnodes == 49 and q.ncells == 9 and q.dofmap.shape == (9, 9)
    j = mesh.jitter(m, 0.2, seed=1)
    for loc, (p, qq) in zip((3, 4, 5), ((1, 2), (0, 2), (0, 1))):
        np.testing.assert_allclose(j.x[j.dofmap[:, loc]], 0.5 * (j.x[j.dofmap[:, p]] + j.x[j.dofmap[:, qq]]), atol=1e-15)
    bc, g = mesh.dirichlet_markers(m)
    assert bc.sum() == 2 * 2 * 7 and g.max() == 0.01
