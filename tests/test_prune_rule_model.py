"""CPU model of the pruning rule of the tensor-core Chamfer filter (csrc/chamfer_prep.cu, chamfer_tc_plan_kernel).

A block (row block r, chunk c) is skipped for the row direction when gap(box_r, box_c)^2 > T_r, where T_r is the max
over the block's rows of an UPPER bound of the row's nearest-target distance (its distance to a few representatives), and
for the column direction when gap^2 > U_c likewise.  The rule is sound if no skipped block contains a row's (column's)
true nearest neighbour or a tie of it - whatever order the clouds are in and whichever representatives are used.
Checked in numpy on primitive-like clouds, in the arrival order and in a Morton-cell order like the prep pass's."""
import numpy as np

BLK = 128


def boxes(p):
    q = p.reshape(-1, BLK, 3)
    return q.min(1), q.max(1)


def gap2(lo_a, hi_a, lo_b, hi_b):
    g = np.maximum(0.0, np.maximum(lo_a[:, None] - hi_b[None], lo_b[None] - hi_a[:, None]))
    return (g * g).sum(-1)


def d2(a, b):
    return ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)


def bounds(mine, other, g_mine_other, near, reps):
    """per block of `mine`: max over its points of the min distance to `reps` representatives of each of the `near`
    blocks of `other` whose boxes are nearest"""
    nb = mine.shape[0] // BLK
    out = np.empty(nb)
    for blk in range(nb):
        sel = np.argsort(g_mine_other[blk], kind="stable")[:near]
        idx = (sel[:, None] * BLK + (np.arange(reps) * (BLK // reps))[None]).ravel()
        out[blk] = d2(mine[blk * BLK:(blk + 1) * BLK], other[idx]).min(1).max()
    return out


def cell_order(p, seg):
    out = np.empty(len(p), np.int64)
    for a in range(0, len(p), seg):
        s = p[a:a + seg]
        lo, hi = s.min(0), s.max(0)
        q = np.minimum(((s - lo) / np.maximum(hi - lo, 1e-30) * 15.999).astype(np.int64), 15)
        code = np.zeros(len(s), np.int64)
        for bit in range(4):
            for k in range(3):
                code |= ((q[:, k] >> bit) & 1) << (3 * bit + k)
        out[a:a + seg] = a + np.argsort(code, kind="stable")
    return out


def scene(rng, k, n, m):
    c = (rng.random((k, 1, 3)) - 0.5) * 0.8
    e = rng.random((k, 1, 3)) * 0.1 + 0.02
    u = rng.random((k, n, 3)) * 2 - 1
    ax = rng.integers(0, 3, (k, n))
    np.put_along_axis(u, ax[..., None], np.sign(rng.random((k, n, 1)) - 0.5), axis=2)
    p1 = (c + u * e).reshape(k * n, 3)
    t = rng.random((m, 3)) * 2 - 1
    ax2 = rng.integers(0, 3, m)
    t[np.arange(m), ax2] = np.sign(rng.random(m) - 0.5)
    p2 = (rng.random((1, 3)) - 0.5) * 0.4 + t * (rng.random((1, 3)) * 0.3 + 0.1)
    return p1, p2


def test_pruned_blocks_never_hold_a_nearest_neighbour():
    rng = np.random.default_rng(3)
    for sort_rows in (False, True):
        for (k, n, m, near_r, reps_r, near_c, reps_c) in ((4, 1024, 1024, 8, 4, 32, 4), (6, 512, 2048, 2, 1, 4, 2)):
            p1, p2 = scene(rng, k, n, m)
            if sort_rows:
                p1 = p1[cell_order(p1, 4096)]
            p2 = p2[cell_order(p2, len(p2))]
            rlo, rhi = boxes(p1); clo, chi = boxes(p2)
            G = gap2(rlo, rhi, clo, chi)                                       # (row blocks, chunks)
            T = bounds(p1, p2, G, near_r, reps_r)
            U = bounds(p2, p1, G.T, near_c, reps_c)
            D = d2(p1, p2)
            skip_r = G > T[:, None] * 1.00001
            skip_c = G > U[None, :] * 1.00001
            # every (row, target) pair at the row's minimum distance (ties included) lies in a block that is kept
            rows, cols = np.nonzero(D <= D.min(1, keepdims=True))
            assert not skip_r[rows // BLK, cols // BLK].any()
            rows, cols = np.nonzero(D <= D.min(0, keepdims=True))
            assert not skip_c[rows // BLK, cols // BLK].any()
            assert skip_r.mean() + skip_c.mean() > 0.1                        # and the rule does prune
