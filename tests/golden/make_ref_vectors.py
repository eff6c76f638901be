"""Generates tests/golden/ref_p1_vectors.json by RUNNING THE REFERENCE'S OWN CODE:
oracle/_ref/libref_{B,blocks}.so = damIntegrator + asym_stress of
MFEM/mechanic2d/asym_elasto_damage_model.cc compiled in place (oracle/ref_shim/
build_ref.sh) against a minimal MFEM stand-in.  Run in the build container only
(the reference tree does not exist on the GPU box):

    make -C oracle ref && python tests/golden/make_ref_vectors.py

Cases (inputs stored with the outputs, floats as hex for bit-exact round trips):
  * every triangle of common/data/square.msh, orientation fixed as MFEM does at load
    (M.cc:1020, fix_orientation: vertices 0 and 1 swapped when clockwise), materials by
    physical tag (M.cc:1086-1098), d = 0: element matrices from BOTH reference code
    paths (USE_B: M.cc:699-704,886-887; blocks: M.cc:705-717,893-911);
  * damaged cases (d > 0) on a subset of those triangles with seeded random
    displacements: closed-form tangent (M.cc:736-872), residual with asym_stress
    (M.cc:207-329) and the load term (M.cc:613-632), incl. the special branches
    (null strain, d = 1 pure traction, |eps_xy| below the limit, r below the limit).
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fem-libraries_b200")]
from oracle import oracle  # noqa: E402  (only for the Young-modulus table / Lame pair)

HERE = os.path.dirname(os.path.abspath(__file__))
dp = C.POINTER(C.c_double)


def load(name):
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", name))
    L.ref_element_grad.argtypes = [dp, C.c_double, C.c_double, C.c_double, dp, dp]
    L.ref_element_vector.argtypes = [dp, C.c_double, C.c_double, C.c_double, dp, dp, dp]
    L.ref_load_points.argtypes = [dp]
    return L


def P(a):
    return a.ctypes.data_as(dp)


def grad(L, xv, lam, mu, d, elfun):
    out = np.zeros(36)
    xv = np.ascontiguousarray(xv, dtype=np.float64)
    u = np.ascontiguousarray(elfun, dtype=np.float64)
    L.ref_element_grad(P(xv), lam, mu, d, P(u), P(out))
    return out


def vect(L, xv, lam, mu, d, elfun, fq):
    out = np.zeros(6)
    xv = np.ascontiguousarray(xv, dtype=np.float64)
    u = np.ascontiguousarray(elfun, dtype=np.float64)
    fq = np.ascontiguousarray(fq, dtype=np.float64)
    L.ref_element_vector(P(xv), lam, mu, d, P(u), P(fq), P(out))
    return out


def hx(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]


def main():
    LB, LK = load("libref_B.so"), load("libref_blocks.so")
    with open(os.path.join(HERE, "square_mesh.json")) as f:
        m = json.load(f)
    x = np.array([[float(a), float(b)] for a, b in m["x"]])
    tri = np.array(m["triangles"])
    tag = np.array(m["triangle_tag"])
    Etab = oracle.E_table()
    pts = np.zeros(6)
    LB.ref_load_points(P(pts))
    out = {"source": "MFEM/mechanic2d/asym_elasto_damage_model.cc lines 1-330,487-953 compiled in place "
                     "(oracle/ref_shim/build_ref.sh); layouts: elmat 6x6 column-major byNODES, elfun/elvect byNODES",
           "load_points": hx(pts), "linear": [], "damaged": []}
    worst = 0.0
    for e in range(tri.shape[0]):
        v = tri[e].copy()
        xv = x[v]
        det = (xv[1, 0] - xv[0, 0]) * (xv[2, 1] - xv[0, 1]) - (xv[2, 0] - xv[0, 0]) * (xv[1, 1] - xv[0, 1])
        if det < 0:  # MFEM fix_orientation for triangles: swap the first two vertices
            v[[0, 1]] = v[[1, 0]]
            xv = x[v]
        lam, mu = oracle.lame(Etab[tag[e] % 200], 0.3)
        kB = grad(LB, xv, lam, mu, 0.0, np.zeros(6))
        kK = grad(LK, xv, lam, mu, 0.0, np.zeros(6))
        worst = max(worst, np.abs(kB - kK).max() / np.abs(kB).max())
        out["linear"].append({"cell": e, "vertices": v.tolist(), "xv": hx(xv), "lam": lam.hex(), "mu": mu.hex(),
                              "elmat_B": hx(kB), "elmat_blocks": hx(kK)})
    print("linear: B vs blocks paths of the reference agree to", worst)

    rng = np.random.default_rng(20261018)
    special = [("null_strain", 0.4, np.zeros(6)), ("full_damage_traction", 1.0, None), ("shear_free", 0.5, None),
               ("isotropic", 0.6, None)]
    cases = [("random", float(rng.uniform(0.02, 0.98)), None) for _ in range(36)] + special
    for k, (kind, d, u) in enumerate(cases):
        e = int(rng.integers(0, tri.shape[0]))
        v = np.array(out["linear"][e]["vertices"])
        xv = x[v]
        lam, mu = oracle.lame(Etab[tag[e] % 200], 0.3)
        if u is None:
            u = 1e-3 * rng.standard_normal(6)
        if kind == "full_damage_traction":      # u = pure dilatation: both eigenvalues positive
            u = np.concatenate([2e-3 * xv[:, 0], 1e-3 * xv[:, 1]])
        elif kind == "shear_free":               # eps_xy = 0 exactly, eps_yy > eps_xx (quirk B1 of SURVEY.md)
            u = np.concatenate([1e-3 * xv[:, 0], 3e-3 * xv[:, 1]])
        elif kind == "isotropic":                # eps_xx = eps_yy, eps_xy = 0: r < limit branch
            u = np.concatenate([-2e-3 * xv[:, 0], -2e-3 * xv[:, 1]])
        fq = 1e5 * rng.standard_normal(6)
        out["damaged"].append({"kind": kind, "cell": e, "xv": hx(xv), "lam": lam.hex(), "mu": mu.hex(), "d": float(d).hex(),
                               "elfun": hx(u), "fq": hx(fq),
                               "elmat_B": hx(grad(LB, xv, lam, mu, d, u)), "elmat_blocks": hx(grad(LK, xv, lam, mu, d, u)),
                               "elvect": hx(vect(LB, xv, lam, mu, d, u, fq)),
                               "elvect_noload": hx(vect(LB, xv, lam, mu, d, u, np.zeros(6)))})
    with open(os.path.join(HERE, "ref_p1_vectors.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out["linear"]), "linear and", len(out["damaged"]), "damaged cases")


if __name__ == "__main__":
    main()
