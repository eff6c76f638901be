"""Field operators either side of the hot path: damage-field smoothing
(M.cc:1258-1315, F.py:160-199), DG0 strain / stress output fields (M.cc:333-430,1551-1563),
Gmsh-2.2 ingestion (M.cc:1017-1020).  CPU part: the oracle against hand-computed answers and the
pinned P1 residual; GPU part: the CUDA kernels against the oracle."""
import os

import numpy as np
import pytest

from oracle import oracle
from femb200 import mesh as fm


def write_msh22(path, x, tri, tri_tag, edges, edge_tag):
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(x))
        for i, p in enumerate(x):
            f.write("%d %.17g %.17g 0\n" % (i + 1, p[0], p[1]))
        f.write("$EndNodes\n$Elements\n%d\n" % (len(tri) + len(edges)))
        k = 1
        for e, t in zip(edges, edge_tag):
            f.write("%d 1 2 %d %d %d %d\n" % (k, t, t, e[0] + 1, e[1] + 1))
            k += 1
        for c, t in zip(tri, tri_tag):
            f.write("%d 2 2 %d %d %d %d %d\n" % (k, t, t, c[0] + 1, c[1] + 1, c[2] + 1))
            k += 1
        f.write("$EndElements\n")


def square_as_mesh(square, tmp_path):
    p = os.path.join(tmp_path, "square.msh")
    write_msh22(p, square["x"], square["tri"], square["tag"], square["edges"], square["edge_tag"])
    return fm.read_gmsh22(p)


# ---------------------------------------------------------------------------
# CPU: reader, oracle
# ---------------------------------------------------------------------------
def test_gmsh_reader_roundtrip(square, tmp_path):
    m = square_as_mesh(square, str(tmp_path))
    assert m.etype == fm.P1 and m.nnodes == 62 and m.ncells == 98
    np.testing.assert_array_equal(m.x, square["x"])
    np.testing.assert_array_equal(m.dofmap, square["tri"])
    np.testing.assert_array_equal(m.meta["cell_tags"], square["tag"])
    np.testing.assert_array_equal(m.meta["facets"], square["edges"])
    np.testing.assert_array_equal(m.meta["facet_tags"], square["edge_tag"])
    E = fm.young_from_tags(m.meta["cell_tags"])
    tab = fm.young_table()
    assert set(np.unique(E)) == {tab[1], tab[2]}                            # physical faces 1 and 2
    d0 = fm.damage_seed(m, [4])                                             # M.cc:1164-1167
    assert d0.sum() == 8.0 and set(np.unique(d0)) == {0.0, 1.0}             # 7 segments of tag 4 -> 8 nodes


def test_smoothing_path_graph_by_hand():
    # 5 vertices in a strip of triangles is awkward to do by hand; use two triangles sharing an
    # edge: vertices 0-1-2 and 1-2-3.  deg = (2, 3, 3, 2).  d0 = (1, 0, 0, 0).
    tri = np.array([[0, 1, 2], [1, 2, 3]], dtype=np.int32)
    d = oracle.smooth_damage(4, tri, np.array([1., 0., 0., 0.]), niter=1)
    # sweep A (only where d < 0.01): s = (-, 1, 1, 0) -> d = (1, 1/3, 1/3, 0)
    # sweep B: s = (2/3, 1 + 1/3, 1 + 1/3, 2/3) -> d = max(s / deg, d) = (1, 4/9, 4/9, 1/3)
    third = 1. * (1. / 3.)
    a = (1. + third) * (1. / 3.)
    np.testing.assert_array_equal(d, [1.0, a, a, (third + third) * 0.5])


def test_smoothing_properties_square(square):
    nv = len(square["x"])
    d0 = np.zeros(nv)
    d0[np.unique(square["edges"][square["edge_tag"] == 4])] = 1.0
    prev = d0
    for it in (1, 2, 8):
        d = oracle.smooth_damage(nv, square["tri"], d0, niter=it)
        assert np.all(d >= prev - 0.0) and np.all(d <= 1.0) and np.all(d[d0 == 1.0] == 1.0)
        assert (d > 0).sum() >= (prev > 0).sum()
        prev = d
    assert (prev > 0).sum() > (d0 > 0).sum()


def test_strain_stress_linear_field_and_pinned_residual(square):
    x, tri = square["x"], square["tri"]
    E = fm.young_table()[square["tag"] % 200]
    A = np.array([[0.01, 0.004], [-0.002, -0.003]])
    u = (x @ A.T).reshape(-1)
    rng = np.random.default_rng(5)
    dn = np.where(rng.random(len(x)) < 0.4, rng.random(len(x)) * 0.9, 0.0)
    eps, sig = oracle.cell_strain_stress(oracle.P1, x, tri, tri, E, 0.3, u, dnod=dn)
    np.testing.assert_allclose(eps, np.tile([A[0, 0], 0.5 * (A[0, 1] + A[1, 0]), A[1, 1]], (len(tri), 1)), rtol=0,
                               atol=1e-15)
    # the stress output and the (reference-pinned) P1 residual are the same asym_stress:
    # r_e = |T| G sigma  (M.cc:586-611)
    for e in range(len(tri)):
        xv = x[tri[e]]
        l, m = oracle.lame(E[e], 0.3)
        d = dn[tri[e]].mean()
        r = oracle.p1_element_vector(xv.reshape(-1), l, m, d, u.reshape(-1, 2)[tri[e]].reshape(-1))
        J = np.array([xv[1] - xv[0], xv[2] - xv[0]]).T
        G = np.array([[-1., -1.], [1., 0.], [0., 1.]]) @ np.linalg.inv(J)
        S = np.array([[sig[e, 0], sig[e, 1]], [sig[e, 1], sig[e, 2]]])
        want = 0.5 * abs(np.linalg.det(J)) * (G @ S)
        np.testing.assert_allclose(np.asarray(r).reshape(3, 2), want, rtol=1e-11, atol=1e-9 * abs(want).max())


# ---------------------------------------------------------------------------
# GPU: CUDA kernels against the oracle
# ---------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_smoothing_square_bit_exact(square, tmp_path):
    from femb200 import fem
    m = square_as_mesh(square, str(tmp_path))
    d0 = fm.damage_seed(m, [4])
    sm = fem.DamageSmoother(m)
    for it in (0, 1, 8):
        got = sm.smooth(d0, niter=it).cpu().numpy()
        np.testing.assert_array_equal(got, oracle.smooth_damage(m.nnodes, m.xdofmap, d0, niter=it))


@pytest.mark.gpu
@pytest.mark.parametrize("order,n", [(1, 37), (2, 21)])
def test_gpu_smoothing_structured_bit_exact(order, n):
    from femb200 import fem
    m = fm.jitter(fm.structured_triangles(n, n + 3, order=order), 0.2, seed=2)
    rng = np.random.default_rng(11)
    d0 = np.zeros(m.nnodes)
    verts = np.unique(m.xdofmap)
    d0[rng.choice(verts, size=max(3, len(verts) // 40), replace=False)] = 1.0
    got = fem.DamageSmoother(m).smooth(d0, niter=8).cpu().numpy()
    want = oracle.smooth_damage(m.nnodes, m.xdofmap, d0, niter=8)
    np.testing.assert_array_equal(got, want)
    assert np.all(got[np.setdiff1d(np.arange(m.nnodes), verts)] == 0.0)    # P2 edge nodes carry no damage


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n", [("P1", 33), ("P2", 19), ("Q2", 14)])
def test_gpu_strain_stress(kind, n):
    from femb200 import fem
    m = fm.structured_triangles(n, n + 2, order=1 if kind == "P1" else 2) if kind != "Q2" else fm.structured_quads_q2(n, n + 2)
    m = fm.jitter(m, 0.2, seed=4)
    E = fm.young_per_cell(m.ncells)
    rng = np.random.default_rng(9)
    u = 1e-3 * rng.standard_normal(m.ndofs)
    dn = fm.damage_band(m)
    form = fem.ElasticityForm(m, E, 0.3, d=dn, u=u)
    eps, sig = fem.cell_strain_stress(form)
    weps, wsig = oracle.cell_strain_stress(m.etype, m.x, m.xdofmap, m.dofmap, E, 0.3, u, dnod=dn)
    assert (dn[m.xdofmap].mean(axis=1) > 0).any()
    np.testing.assert_allclose(eps.cpu().numpy(), weps, rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(sig.cpu().numpy(), wsig, rtol=1e-11, atol=1e-9 * np.abs(wsig).max())


@pytest.mark.gpu
@pytest.mark.parametrize("refine", [0, 1])
def test_gpu_reference_driver_sequence(square, tmp_path, refine):
    """BASELINE config 1 end to end (refine = 1: with one level of the -r / MAX_REFINE loop, 8 (r + 1) smoothing sweeps): the reference driver's stages (mesh file, materials by tag, damage seed +
    smoothing, BCs, load, Newton, strain/stress) through examples/mechanic2d_square.py against the same
    sequence written with the oracle."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("mechanic2d_square", os.path.join(root, "examples", "mechanic2d_square.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    p = os.path.join(str(tmp_path), "square.msh")
    ex.fixture_msh(p)
    out = ex.run(p, max_refine=refine, verbose=False)
    m = out["mesh"]
    np.testing.assert_array_equal(m.x[:len(square["x"])], square["x"])
    assert m.ncells == len(square["tri"]) * 4 ** refine
    d = oracle.smooth_damage(m.nnodes, m.xdofmap, out["d0"], niter=8 * (refine + 1))
    np.testing.assert_array_equal(out["d"].cpu().numpy(), d)
    assert (d > 0).sum() > (out["d0"] > 0).sum() and d.max() == 1.0 and d.min() >= 0.0
    want, it_o, norms_o = oracle.newton(m.etype, m.x, m.xdofmap, m.dofmap, out["E"], 0.3, out["bc"], out["g"], dnod=d,
                                        fnod=out["load"].ravel())
    u = out["u"].cpu().numpy()
    assert out["newton"].iterations == it_o
    assert np.linalg.norm(u - want) / np.linalg.norm(want) < 1e-9
    eps, sig = oracle.cell_strain_stress(m.etype, m.x, m.xdofmap, m.dofmap, out["E"], 0.3, u, dnod=d)
    np.testing.assert_allclose(out["strain"].cpu().numpy(), eps, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(out["stress"].cpu().numpy(), sig, rtol=1e-9, atol=1e-9 * np.abs(sig).max())


def test_comparison_harness_roundtrip(square, tmp_path):
    """OUT_COMP / IN_COMP files of the reference (M.cc:1660-1725, F.cc:1036-1130): write, read back, and the
    two matching rules give the per-component L2 norms of a known perturbation."""
    from femb200 import compare
    x = square["x"]
    rng = np.random.default_rng(1)
    u = 1e-2 * rng.standard_normal((len(x), 2))
    p = os.path.join(str(tmp_path), "mfem_disp_0")
    compare.write_disp_file(p, x, u.ravel())
    assert os.path.getsize(p) == 32 * len(x)                                 # four raw doubles per vertex
    xr, ur = compare.read_disp_file(p)
    np.testing.assert_array_equal(xr, x)
    np.testing.assert_array_equal(ur, u)
    du = 1e-6 * rng.standard_normal(u.shape)
    want = (np.sqrt((du[:, 0] ** 2).sum()), np.sqrt((du[:, 1] ** 2).sum()))
    got = compare.compare_disp_file(p, x, (u + du).ravel())
    np.testing.assert_allclose(got, want, rtol=1e-12)
    perm = rng.permutation(len(x))                                           # FEniCSx rule: other dof order
    got = compare.compare_disp_file(p, x[perm] * (1 + 1e-9), (u + du)[perm].ravel(), match="coords")
    np.testing.assert_allclose(got, want, rtol=1e-9)
    with pytest.raises(ValueError, match="vertex"):
        compare.compare_disp_file(p, x + 1e-3, u.ravel())
    with pytest.raises(ValueError, match="no vertex"):
        compare.compare_disp_file(p, x + 1e-3, u.ravel(), match="coords")


def test_gmsh_reader_errors_and_sparse_ids(tmp_path):
    """Reader edge cases: non-contiguous node ids, elements of other types ignored, no triangles, wrong format."""
    p = os.path.join(str(tmp_path), "a.msh")
    with open(p, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n4\n10 0 0 0\n20 1 0 0\n30 0 1 0\n45 1 1 0\n$EndNodes\n"
                "$Elements\n4\n1 15 2 9 9 10\n2 1 2 7 7 10 20\n3 2 2 3 3 10 20 30\n4 2 2 5 5 20 45 30\n$EndElements\n")
    m = fm.read_gmsh22(p)
    assert m.nnodes == 4 and m.ncells == 2
    np.testing.assert_array_equal(m.dofmap, [[0, 1, 2], [1, 3, 2]])
    np.testing.assert_array_equal(m.meta["cell_tags"], [3, 5])
    np.testing.assert_array_equal(m.meta["facets"], [[0, 1]])
    np.testing.assert_array_equal(fm.damage_seed(m, [7]), [1., 1., 0., 0.])
    np.testing.assert_array_equal(fm.damage_seed(m, [8]), [0., 0., 0., 0.])
    tab = fm.young_table()
    np.testing.assert_array_equal(fm.young_from_tags(np.array([3, 205])), [tab[3], tab[5]])
    q = os.path.join(str(tmp_path), "b.msh")
    with open(q, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n2\n1 0 0 0\n2 1 0 0\n$EndNodes\n$Elements\n1\n1 1 2 1 1 1 2\n$EndElements\n")
    with pytest.raises(ValueError, match="no triangles"):
        fm.read_gmsh22(q)
    u = os.path.join(str(tmp_path), "u.msh")   # an element names a node that $Nodes does not define
    with open(u, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n3\n1 0 0 0\n2 1 0 0\n4 0 1 0\n$EndNodes\n$Elements\n1\n1 2 2 1 1 1 2 3\n$EndElements\n")
    with pytest.raises(ValueError, match="does not define"):
        fm.read_gmsh22(u)
    r = os.path.join(str(tmp_path), "c.msh")
    with open(r, "w") as f:
        f.write("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$Nodes\n0\n$EndNodes\n$Elements\n0\n$EndElements\n")
    with pytest.raises(ValueError, match="2.x"):
        fm.read_gmsh22(r)


def test_xdmf_roundtrip_and_errors(square, tmp_path):
    """XDMF ingestion of the FEniCSx driver (F.cc:155-163; gmsh_to_xdmf_neper_dam.py): the reference's mesh written in
    the mesh + `_cells` + `_facets` grid layout and read back (geometry bit for bit, cell tags matched by vertex
    triple even when the tag grid lists the cells in another order and orientation), HDF5 data items refused with a
    message when h5py is absent, unknown grid / topology errors."""
    m = square_as_mesh(square, str(tmp_path))
    p = os.path.join(str(tmp_path), "square.xdmf")
    fm.write_xdmf(p, m, "neper_dam")
    r = fm.read_xdmf(p, "neper_dam")
    assert r.etype == fm.P1 and fm.read_xdmf(p).ncells == m.ncells          # default: first grid
    np.testing.assert_array_equal(r.x, m.x)
    np.testing.assert_array_equal(r.xdofmap, m.xdofmap)
    np.testing.assert_array_equal(r.meta["cell_tags"], m.meta["cell_tags"])
    np.testing.assert_array_equal(r.meta["facets"], m.meta["facets"])
    np.testing.assert_array_equal(r.meta["facet_tags"], m.meta["facet_tags"])
    np.testing.assert_array_equal(fm.young_from_tags(r.meta["cell_tags"]), fm.young_from_tags(m.meta["cell_tags"]))
    np.testing.assert_array_equal(fm.damage_seed(r, [4]), fm.damage_seed(m, [4]))
    # tag grid in another cell order, vertices rotated: still matched to the mesh cells
    rng = np.random.default_rng(1)
    perm = rng.permutation(m.ncells)
    m2 = fm.Mesh(fm.P1, m.x, m.xdofmap, m.dofmap, 0, 0, dict(m.meta))
    q = os.path.join(str(tmp_path), "shuffled.xdmf")
    fm.write_xdmf(q, m2, "g")
    txt = open(q).read()
    import xml.etree.ElementTree as ET
    root = ET.fromstring(txt)
    tg = [g for g in root.iter("Grid") if g.get("Name") == "g_cells"][0]
    tg.find("Topology").find("DataItem").text = "\n".join(" ".join(str(v) for v in np.roll(row, 1)) for row in m.xdofmap[perm])
    tg.find("Attribute").find("DataItem").text = "\n".join(str(v) for v in m.meta["cell_tags"][perm])
    ET.ElementTree(root).write(q)
    np.testing.assert_array_equal(fm.read_xdmf(q, "g").meta["cell_tags"], m.meta["cell_tags"])
    # the geometry in a raw binary side file (XDMF "Binary" data item, big-endian float64 after a 16-byte header)
    bx = os.path.join(str(tmp_path), "geom.bin")
    with open(bx, "wb") as fb:
        fb.write(b"\0" * 16)
        fb.write(m.x.astype(">f8").tobytes())
    root = ET.fromstring(open(p).read())
    g0 = [g for g in root.iter("Grid") if g.get("Name") == "neper_dam"][0]
    di = g0.find("Geometry").find("DataItem")
    di.set("Format", "Binary"), di.set("Endian", "Big"), di.set("Precision", "8"), di.set("Seek", "16")
    di.text = "geom.bin"
    pb = os.path.join(str(tmp_path), "square_bin.xdmf")
    ET.ElementTree(root).write(pb)
    np.testing.assert_array_equal(fm.read_xdmf(pb, "neper_dam").x, m.x)
    with pytest.raises(ValueError, match="no grid named"):
        fm.read_xdmf(p, "other")
    h = os.path.join(str(tmp_path), "h.xdmf")
    with open(h, "w") as f:
        f.write('<Xdmf><Domain><Grid Name="a"><Topology TopologyType="Triangle"><DataItem Dimensions="1 3" Format="HDF">'
                'a.h5:/Mesh/a/topology</DataItem></Topology><Geometry GeometryType="XY"><DataItem Dimensions="3 2" '
                'Format="XML">0 0 1 0 0 1</DataItem></Geometry></Grid></Domain></Xdmf>')
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ValueError, match="h5py"):
            fm.read_xdmf(h)
    t = os.path.join(str(tmp_path), "t.xdmf")
    with open(t, "w") as f:
        f.write('<Xdmf><Domain><Grid Name="a"><Topology TopologyType="Quadrilateral"><DataItem Dimensions="1 4" '
                'Format="XML">0 1 2 3</DataItem></Topology><Geometry GeometryType="XY"><DataItem Dimensions="4 2" '
                'Format="XML">0 0 1 0 1 1 0 1</DataItem></Geometry></Grid></Domain></Xdmf>')
    with pytest.raises(ValueError, match="Triangle"):
        fm.read_xdmf(t)
    b = os.path.join(str(tmp_path), "bad.xdmf")
    with open(b, "w") as f:
        f.write('<Xdmf><Domain><Grid Name="a"><Topology TopologyType="Triangle"><DataItem Dimensions="1 3" '
                'Format="XML">0 1 7</DataItem></Topology><Geometry GeometryType="XYZ"><DataItem Dimensions="3 3" '
                'Format="XML">0 0 0 1 0 0 0 1 0</DataItem></Geometry></Grid></Domain></Xdmf>')
    with pytest.raises(ValueError, match="outside the geometry"):
        fm.read_xdmf(b)


def _areas(m):
    t = m.xdofmap.astype(np.int64)
    a, b, c = m.x[t[:, 0]], m.x[t[:, 1]], m.x[t[:, 2]]
    return 0.5 * ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1]))


def test_refine_uniform_square_mesh(square, tmp_path):
    """The -r / MAX_REFINE stage (M.cc:1037-1038, F.cc:166-185) on the reference's mesh: 4 children per cell, one new
    vertex per edge, conforming, same area, orientation kept, cell and facet tags inherited; two levels."""
    m0 = square_as_mesh(square, str(tmp_path))
    area0 = _areas(m0)
    m1 = fm.refine_uniform(m0)
    tri = m0.xdofmap.astype(np.int64)
    edges = np.unique(np.sort(np.stack([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]], 1).reshape(-1, 2), axis=1), axis=0)
    assert m1.ncells == 4 * m0.ncells and m1.nnodes == m0.nnodes + len(edges)
    a1 = _areas(m1)
    np.testing.assert_allclose(a1.reshape(-1, 4).sum(1), area0, rtol=1e-13)
    np.testing.assert_allclose(a1, np.repeat(area0 / 4, 4), rtol=1e-12)          # signed: orientation kept
    np.testing.assert_array_equal(m1.meta["cell_tags"], np.repeat(m0.meta["cell_tags"], 4))
    # conformity: every edge belongs to one (boundary) or two cells, and the tagged facets (boundary or interior
    # lines) are edges of the refined triangulation
    t1 = m1.xdofmap.astype(np.int64)
    e1 = np.sort(np.stack([t1[:, [0, 1]], t1[:, [1, 2]], t1[:, [2, 0]]], 1).reshape(-1, 2), axis=1)
    ue, cnt = np.unique(e1, axis=0, return_counts=True)
    assert set(cnt.tolist()) <= {1, 2}
    alle = {tuple(e) for e in ue.tolist()}
    f1 = np.sort(m1.meta["facets"].astype(np.int64), axis=1)
    assert len(f1) == 2 * len(m0.meta["facets"]) and all(tuple(e) in alle for e in f1.tolist())
    np.testing.assert_array_equal(m1.meta["facet_tags"], np.repeat(m0.meta["facet_tags"], 2))
    # the two halves of a facet meet at the midpoint of the parent facet
    f0 = m0.meta["facets"].astype(np.int64)
    np.testing.assert_allclose(m1.x[m1.meta["facets"][0::2, 1]], 0.5 * (m0.x[f0[:, 0]] + m0.x[f0[:, 1]]), rtol=0, atol=1e-15)
    # the seeded damage of the refined mesh covers the seeded nodes of the coarse one plus the facet midpoints
    d0, d1 = fm.damage_seed(m0, [4]), fm.damage_seed(m1, [4])
    assert np.all(d1[:m0.nnodes] == d0) and d1.sum() == d0.sum() + (m0.meta["facet_tags"] == 4).sum()
    m2 = fm.refine_uniform(m0, 2)
    assert m2.ncells == 16 * m0.ncells and m2.meta["refined"] == 2
    np.testing.assert_allclose(_areas(m2).sum(), area0.sum(), rtol=1e-13)
    with pytest.raises(ValueError, match="P1"):
        fm.refine_uniform(fm.structured_triangles(2, order=2))


def test_refined_mesh_operator_properties(square, tmp_path):
    """The oracle's tangent on the refined reference mesh: rigid-body modes in the null space, symmetric."""
    m = fm.refine_uniform(square_as_mesh(square, str(tmp_path)))
    E = fm.young_from_tags(m.meta["cell_tags"])
    rp, ci = oracle.build_pattern(m.nnodes, m.dofmap)
    K = oracle.assemble_matrix(m.etype, m.x, m.xdofmap, m.dofmap, E, 0.3, rp, ci)
    import scipy.sparse as sp
    A = sp.csr_matrix((K, ci, rp), shape=(m.ndofs, m.ndofs))
    fro = np.sqrt((K * K).sum())
    x, y = m.x[:, 0], m.x[:, 1]
    for mode in (np.stack([np.ones_like(x), 0 * x], 1), np.stack([0 * x, np.ones_like(x)], 1), np.stack([-y, x], 1)):
        assert np.abs(A @ mode.ravel()).max() <= 1e-12 * fro
    assert abs(A - A.T).max() <= 1e-12 * fro
