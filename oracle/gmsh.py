"""Minimal Gmsh-2.2 ASCII reader (points, lines, triangles) for the reference's
only shipped mesh, common/data/square.msh (role of `Mesh(mesh_file,1,0,true)`,
M.cc:1020, without the orientation fix: SURVEY.md B6).  Oracle-side only."""
from __future__ import annotations

import numpy as np


def read_msh22(path: str) -> dict:
    with open(path) as f:
        lines = [ln.strip() for ln in f]
    it = iter(range(len(lines)))
    out = {}
    i = 0
    while i < len(lines):
        if lines[i] == "$Nodes":
            n = int(lines[i + 1])
            rows = [lines[i + 2 + k].split() for k in range(n)]
            ids = np.array([int(r[0]) for r in rows])
            assert np.array_equal(ids, np.arange(1, n + 1)), "non-contiguous node ids"
            out["x"] = np.array([[float(r[1]), float(r[2])] for r in rows], dtype=np.float64)
            i += 2 + n
        elif lines[i] == "$Elements":
            n = int(lines[i + 1])
            tris, ttag, edges, etag = [], [], [], []
            for k in range(n):
                p = [int(t) for t in lines[i + 2 + k].split()]
                etype, ntags = p[1], p[2]
                tags, nodes = p[3:3 + ntags], p[3 + ntags:]
                if etype == 2:
                    tris.append([v - 1 for v in nodes]); ttag.append(tags[0])
                elif etype == 1:
                    edges.append([v - 1 for v in nodes]); etag.append(tags[0])
            out["triangles"] = np.array(tris, dtype=np.int32)
            out["triangle_tag"] = np.array(ttag, dtype=np.int32)
            out["edges"] = np.array(edges, dtype=np.int32)
            out["edge_tag"] = np.array(etag, dtype=np.int32)
            i += 2 + n
        else:
            i += 1
    return out
