"""Generates tests/golden/square_mesh.json and square_kat.json from the
reference's only shipped mesh (/root/reference/common/data/square.msh).

Run in the build container only (the reference tree does not exist on the GPU
box).  The mesh fixture holds topology + coordinates + physical tags; the KAT
file holds the known answers derived from the reference's formulas on that mesh
(SURVEY.md 8c), recomputed here with the oracle and cross-checked against the
values recorded in SURVEY.md.

    python tests/golden/make_square_fixture.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gmsh, oracle  # noqa: E402

REF = "/root/reference/common/data/square.msh"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    m = gmsh.read_msh22(REF)
    x, tri, tag = m["x"], m["triangles"], m["triangle_tag"]
    fixture = {
        "source": "common/data/square.msh (Gmsh 2.2, Neper), topology and coordinates only",
        "x": [[repr(float(a)), repr(float(b))] for a, b in x],
        "triangles": tri.tolist(),
        "triangle_tag": tag.tolist(),
        "edges": m["edges"].tolist(),
        "edge_tag": m["edge_tag"].tolist(),
    }
    with open(os.path.join(HERE, "square_mesh.json"), "w") as f:
        json.dump(fixture, f)

    # --- known answers (P1, d = 0, no BCs), SURVEY.md 8c -------------------
    Etab = oracle.E_table()
    E = Etab[tag % 200]                       # F.cc:545 (tag % 200); M.cc:1093 ((i+1) % 200, attr = i+1)
    nn = x.shape[0]
    rowptr, colidx = oracle.build_pattern(nn, tri)
    vals = oracle.assemble_matrix(oracle.P1, x, tri, tri, E, 0.3, rowptr, colidx)
    import scipy.sparse as sp
    K = sp.csr_matrix((vals, colidx, rowptr), shape=(2 * nn, 2 * nn))
    u = np.zeros(2 * nn)
    u[0::2] = 0.01 * x[:, 0]
    u[1::2] = -0.003 * x[:, 1]
    area = 0.5 * np.abs((x[tri[:, 1], 0] - x[tri[:, 0], 0]) * (x[tri[:, 2], 1] - x[tri[:, 0], 1])
                        - (x[tri[:, 2], 0] - x[tri[:, 0], 0]) * (x[tri[:, 1], 1] - x[tri[:, 0], 1]))
    lam1, mu1 = oracle.lame(Etab[1], 0.3)
    lam2, mu2 = oracle.lame(Etab[2], 0.3)
    kat = {
        "nodes": int(nn), "triangles": int(tri.shape[0]),
        "ndofs": int(2 * nn), "nnz": int(rowptr[-1]), "node_blocks": int(rowptr[-1] // 4),
        "E_range_1": float(Etab[1]), "E_range_2": float(Etab[2]),
        "lame_phys1": [lam1, mu1], "lame_phys2": [lam2, mu2],
        "area_phys1": float(area[tag == 1].sum()), "area_phys2": float(area[tag == 2].sum()),
        "ntri_phys1": int((tag == 1).sum()), "ntri_phys2": int((tag == 2).sum()),
        "fro_norm": float(np.sqrt((vals ** 2).sum())),
        "trace": float(K.diagonal().sum()),
        "energy_uKu": float(u @ (K @ u)),
        "all_clockwise": bool(np.all((x[tri[:, 1], 0] - x[tri[:, 0], 0]) * (x[tri[:, 2], 1] - x[tri[:, 0], 1])
                                     - (x[tri[:, 2], 0] - x[tri[:, 0], 0]) * (x[tri[:, 1], 1] - x[tri[:, 0], 1]) < 0)),
        "survey_values": {   # as recorded by the survey, independent derivation
            "fro_norm": 2.053300217555554e+09, "trace": 1.698801350025497e+10,
            "energy_uKu": 6.255396994989912e+03, "energy_closed_form": 6.255396994989910e+03,
            "nnz": 1520, "node_blocks": 380, "E_range_1": 70402010.05025125, "E_range_2": 26005025.12562814,
            "lame_phys1": [40616544.25976033, 27077696.173173554],
            "lame_phys2": [15002899.11093931, 10001932.740626207],
            "area_phys1": 0.6709746943275, "area_phys2": 0.3290253056725,
        },
    }
    with open(os.path.join(HERE, "square_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    for k in ("fro_norm", "trace", "energy_uKu", "nnz", "area_phys1"):
        print(k, kat[k], kat["survey_values"][k])


if __name__ == "__main__":
    main()
