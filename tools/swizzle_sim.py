"""Developer tool (CPU only): shared-memory wavefronts of the staging stores of assemble_fast_kernel on a
structured P2 mesh, for candidate swizzle keys.

Emulates the plan (pattern, rotational visit order, flips and carries of k_fast_records, ranking inside 64-node
tiles) in numpy, lists the 16-byte staging units every STS.128 of the kernel touches (warp = 16 ranks x 2 scalar
rows, four quarter-warp phases of 8 lanes), and counts wavefronts = sum over phases of the largest number of
distinct units that share a 16-byte bank group.  The swizzle is u ^ key(u >> 3) with a 3-bit key.

    python tools/swizzle_sim.py [n]          # n x n cells, default 160 (n = 96: about 25 minutes for the full sweep)
"""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "fem-libraries_b200")]
from femb200 import mesh as fm  # noqa: E402

R = 64


def build(n):
    m = fm.structured_triangles(n, order=2)
    dm = m.dofmap.astype(np.int64)
    nn, nc = m.nnodes, m.ncells
    # node -> (cell, a) visits, ascending cell
    node = dm.ravel()
    order = np.argsort(node, kind="stable")
    cell, a = order // 6, order % 6
    nptr = np.zeros(nn + 1, dtype=np.int64)
    np.add.at(nptr, node + 1, 1)
    nptr = np.cumsum(nptr)
    # pattern: sorted unique neighbours
    nb = [None] * nn
    for I in range(nn):
        cs = cell[nptr[I]:nptr[I + 1]]
        nb[I] = np.unique(dm[cs].ravel())
    deg = np.array([len(x) for x in nb])
    brp = np.concatenate([[0], np.cumsum(deg)])
    return m, dm, cell, a, nptr, nb, deg, brp


def rotated(dmrow, aa):
    m_ = aa - 3 if aa >= 3 else aa
    v = [dmrow[(m_ + t) % 3] for t in range(3)]
    e = [dmrow[3 + (m_ + t) % 3] for t in range(3)]
    return v + e      # positions 0..5


def accesses(n):
    m, dm, cell, a, nptr, nb, deg, brp = build(n)
    nn = m.nnodes
    out = []          # (tile, level, warp, put_seq, lane, unit)
    for t0 in range(0, nn, R):
        nodes = np.arange(t0, min(t0 + R, nn))
        cnts = nptr[nodes + 1] - nptr[nodes]
        rank_order = sorted(range(len(nodes)), key=lambda i: (-cnts[i], i))
        b0 = brp[t0]
        for rank, i in enumerate(rank_order):
            I = nodes[i]
            vis = [(cell[k], a[k]) for k in range(nptr[I], nptr[I + 1])]
            if not vis:
                continue
            vert = vis[0][1] < 3
            if vert and len(vis) > 1:      # fan walk (k_order_visits)
                p1 = [dm[c][(aa + 1) % 3] for c, aa in vis]
                p2 = [dm[c][(aa + 2) % 3] for c, aa in vis]
                occ = lambda v: sum((x == v) + (y == v) for x, y in zip(p1, p2))
                cur, entry = -1, 0
                for pas in (0, 1):
                    if cur >= 0:
                        break
                    for k in range(len(vis)):
                        for v in (p1[k], p2[k]):
                            if pas == 0 and occ(v) != 1:
                                continue
                            if cur < 0 or v < entry:
                                cur, entry = k, v
                used, seq = set(), []
                for _ in range(len(vis)):
                    seq.append(cur)
                    used.add(cur)
                    ex = p2[cur] if entry == p1[cur] else p1[cur]
                    nxt = next((j for j in range(len(vis)) if j not in used and ex in (p1[j], p2[j])), -1)
                    entry = ex
                    if nxt < 0:
                        nxt = next((j for j in range(len(vis)) if j not in used), -1)
                        if nxt >= 0:
                            entry = p1[nxt]
                    cur = nxt
                vis = [vis[k] for k in seq]
            slot = {J: s for s, J in enumerate(nb[I])}
            r0 = 2 * (brp[I] - b0)
            cin, carry_v = False, -1
            for j, (c, aa) in enumerate(vis):
                pos = [slot[J] for J in rotated(dm[c], aa)]
                nxtpos = [slot[J] for J in rotated(dm[vis[j + 1][0]], vis[j + 1][1])] if j + 1 < len(vis) else None
                flip, cout = 0, False
                if vert:
                    def side_match(rs):
                        return nxtpos is not None and any(pos[1 + rs] == nxtpos[1 + qs] and pos[5 - rs] == nxtpos[5 - qs]
                                                          for qs in (0, 1))
                    if cin:
                        flip = 0 if pos[1] == carry_v else 1
                        cout = side_match(1 - flip)
                    elif side_match(1):
                        cout = True
                    elif side_match(0):
                        cout, flip = True, 1
                else:
                    if cin:
                        flip = 0 if pos[1] == carry_v else 1
                    else:
                        cout = nxtpos is not None and {pos[1], pos[2]} == {nxtpos[1], nxtpos[2]}
                sl = list(pos)
                if flip:
                    sl[1], sl[2], sl[4], sl[5] = pos[2], pos[1], pos[5], pos[4]
                puts = ([1, 5, 3] + ([] if cout else [2, 4])) if vert else ([0, 4, 5] + ([] if cout else [1, 2]))
                for h in (0, 1):
                    lane = (rank % 16) + 16 * h
                    for seqno, t in enumerate(puts):
                        out.append((t0 // R, j, rank // 16, seqno + (0 if vert else 8), lane, r0 + h * deg[I] + sl[t]))
                cin = cout
                if cout:
                    carry_v = sl[2] if vert else sl[1]
    return np.array(out, dtype=np.int64)


def wavefronts(acc, key_of_g):
    u = acc[:, 5]
    us = u ^ key_of_g[(u >> 3) % len(key_of_g)]
    bank = us & 7
    # one STS instruction = (tile, level, warp, put); phase = lane // 8
    inst = ((acc[:, 0] * 16 + acc[:, 1]) * 8 + acc[:, 2]) * 16 + acc[:, 3]
    ph = inst * 4 + acc[:, 4] // 8
    _, phid = np.unique(ph, return_inverse=True)
    mult = np.zeros((phid.max() + 1, 8), dtype=np.int32)
    np.add.at(mult, (phid, bank), 1)
    return int(mult.max(axis=1).sum()), int(phid.max() + 1)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    acc = accesses(n)
    bitrev = np.array([0, 4, 2, 6, 1, 5, 3, 7])
    ident = np.arange(8)
    w0, ideal = wavefronts(acc, np.zeros(8, dtype=np.int64))
    print(f"n = {n}: {len(acc)} unit stores in {ideal} quarter-warp phases (ideal wavefronts)")
    print(f"no swizzle        : {w0} wavefronts ({w0 / ideal:.3f} x ideal)")
    for name, key in (("key = g & 7", ident), ("key = bitrev3(g) (shipped)", bitrev)):
        w, _ = wavefronts(acc, key)
        print(f"{name:18s}: {w} wavefronts ({w / ideal:.3f} x ideal)")
    best = []
    for perm in itertools.permutations(range(8)):
        if perm[0] != 0:
            continue              # key(0) = 0 without loss of generality (XOR by a constant is a relabelling)
        w, _ = wavefronts(acc, np.array(perm))
        best.append((w, perm))
    best.sort()
    print("best 3-bit tables key[g & 7]:")
    for w, perm in best[:5]:
        print(f"   {perm}: {w} ({w / ideal:.3f} x ideal)")
    # keys that also look at the next three bits of g
    rng = np.random.default_rng(0)
    best64 = (10 ** 18, None)
    for _ in range(3000):
        key = rng.integers(0, 8, size=64)
        w, _ = wavefronts(acc, key)
        if w < best64[0]:
            best64 = (w, key.copy())
    print(f"best of 3000 random 64-entry tables key[g & 63]: {best64[0]} ({best64[0] / ideal:.3f} x ideal)")
