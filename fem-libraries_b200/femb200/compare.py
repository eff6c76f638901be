"""The reference's cross-library comparison harness (its only verification mechanism).

`OUT_COMP` builds of the MFEM driver dump the converged displacement as raw doubles, one
(x, y, ux, uy) quadruple per mesh vertex (`mfem_disp_<refine>`, M.cc:1660-1687); `IN_COMP` builds read
such a file back and print the per-component L2 norms of the difference, the MFEM driver assuming the
same vertex order and asserting the coordinates to 1e-6 (M.cc:1689-1725), the FEniCSx driver matching
the dofs by coordinates to 1e-5 (F.cc:1036-1130).  This module reads and writes that format and
computes the same two numbers for a displacement computed here, so that a `data/mfem_disp_0` produced
by the reference pins the GPU solution without any further glue.  Host side only (numpy).
"""
from __future__ import annotations

import numpy as np


def write_disp_file(path: str, x: np.ndarray, u: np.ndarray) -> None:
    """OUT_COMP format (M.cc:1676-1687): per vertex x, y, ux, uy as raw float64, vertex order."""
    x = np.asarray(x, dtype=np.float64)[:, :2]
    u = np.asarray(u, dtype=np.float64).reshape(-1, 2)
    if u.shape[0] != x.shape[0]:
        raise ValueError(f"{u.shape[0]} displacement pairs for {x.shape[0]} vertices")
    np.ascontiguousarray(np.hstack([x, u])).tofile(path)


def read_disp_file(path: str):
    """(x (n, 2), u (n, 2)) of an OUT_COMP file."""
    raw = np.fromfile(path, dtype=np.float64)
    if raw.size % 4:
        raise ValueError(f"{path}: {raw.size} doubles is not a whole number of (x, y, ux, uy) records")
    q = raw.reshape(-1, 4)
    return q[:, :2].copy(), q[:, 2:].copy()


def compare_disp_file(path: str, x: np.ndarray, u: np.ndarray, match: str = "order", tol: float | None = None):
    """(L2x, L2y) = sqrt(sum (u - u_ref)^2) per component against an OUT_COMP file.

    match = "order":  the MFEM IN_COMP rule: same vertex order, coordinates asserted to `tol` (1e-6).
    match = "coords": the FEniCSx IN_COMP rule: every vertex is looked up by its coordinates
                      (relative tolerance `tol`, 1e-5); a vertex without a partner is an error.
    """
    xr, ur = read_disp_file(path)
    x = np.asarray(x, dtype=np.float64)[:, :2]
    u = np.asarray(u, dtype=np.float64).reshape(-1, 2)
    if match == "order":
        tol = 1e-6 if tol is None else tol
        if xr.shape != x.shape:
            raise ValueError(f"{path}: {xr.shape[0]} vertices in the file, {x.shape[0]} in the mesh")
        bad = np.nonzero(np.abs(xr - x).max(axis=1) >= tol)[0]
        if bad.size:
            raise ValueError(f"{path}: vertex {bad[0]} is at {xr[bad[0]]} in the file and {x[bad[0]]} in the mesh")
        diff = u - ur
    elif match == "coords":
        tol = 1e-5 if tol is None else tol
        scale = np.maximum(np.abs(xr).max(axis=0), 1e-300)
        key = np.round(xr / (tol * scale)).astype(np.int64)           # bucket the file's vertices
        order = np.lexsort((key[:, 1], key[:, 0]))
        ks = key[order]
        diff = np.empty_like(u)
        mine = np.round(x / (tol * scale)).astype(np.int64)
        for i in range(x.shape[0]):                                    # neighbouring buckets cover rounding
            found = -1
            for dx in (0, -1, 1):
                for dy in (0, -1, 1):
                    k0, k1 = mine[i, 0] + dx, mine[i, 1] + dy
                    lo = np.searchsorted(ks[:, 0], k0, "left")
                    hi = np.searchsorted(ks[:, 0], k0, "right")
                    j = lo + np.searchsorted(ks[lo:hi, 1], k1, "left")
                    if j < hi and ks[j, 1] == k1:
                        found = order[j]
                        break
                if found >= 0:
                    break
            if found < 0:
                raise ValueError(f"{path}: no vertex of the file at {x[i]}")
            diff[i] = u[i] - ur[found]
    else:
        raise ValueError("match must be 'order' or 'coords'")
    return float(np.sqrt((diff[:, 0] ** 2).sum())), float(np.sqrt((diff[:, 1] ** 2).sum()))
