"""CPU restatements (numpy) of two algorithmic ideas of the round-2 kernels, checked against the oracle / the reference's formulas:

1. csrc/lgr_bin.cuh -- three stable 11/11/10-bit counting-sort passes on the depth keys, then ONE stable counting-sort pass over the tile
   index in which every block owns a contiguous chunk of the depth-ordered Gaussians (count matrix M[block][bin] -> exclusive scan in
   bin-major order -> in-order scatter).  Must reproduce the order of the reference's stable 64-bit sort on (tile << 32 | depth bits),
   i.e. the oracle's lists (oracle/lgo.c, following rasterizer_impl.cu:70-138).
2. csrc/lgr_blend.cuh -- the backward's single running scalar D: T*(c.dpix) - D/(1-alpha) equals the reference's
   T*sum_c (c - accum_rec)*dpix - T_final/(1-alpha)*bg.dpix with its three colour recurrences (backward.cu:505-518)."""
import numpy as np

from oracle.lgo import Oracle
from tests.util import make_config


def _stable_pass(keys, ids, shift, bits, blocks):
    """one LSD pass as the kernels do it: per-block histograms -> exclusive scan over (digit, block) -> in-order scatter"""
    n, nb = len(keys), 1 << bits
    digit = (keys >> shift) & (nb - 1)
    per = -(-n // blocks)
    M = np.zeros((blocks, nb), np.int64)
    for b in range(blocks):
        d = digit[b * per:(b + 1) * per]
        if len(d):
            M[b] = np.bincount(d, minlength=nb)
    excl = np.cumsum(M.T.reshape(-1)) - M.T.reshape(-1)       # bin-major: all blocks of digit 0, then digit 1, ...
    start = excl.reshape(nb, blocks).T.copy()                  # start[b][digit]
    out_k, out_i = np.empty_like(keys), np.empty_like(ids)
    for b in range(blocks):
        for k in range(b * per, min(n, (b + 1) * per)):       # the block walks its chunk IN ORDER
            pos = start[b, digit[k]]
            start[b, digit[k]] += 1
            out_k[pos], out_i[pos] = keys[k], ids[k]
    return out_k, out_i


def test_block_chunked_counting_sorts_reproduce_the_oracle_lists():
    act, view, _ = make_config("deg1")
    o = Oracle()
    g = o.preprocess(view, act["means3D"], act["opacities"], shs=act["shs"], scales=act["scales"], rotations=act["rotations"])
    want_list, want_ranges = o.bin(view, g["means2D"], g["depths"], g["radii"], g["tiles_touched"])
    P = len(g["radii"])
    gx, gy = (view.W + 15) // 16, (view.H + 15) // 16
    # depth keys as the preprocess kernel writes them: float bits, culled Gaussians last
    keys = np.where(g["radii"] > 0, g["depths"].astype(np.float32).view(np.uint32), np.uint32(0xFFFFFFFF)).astype(np.uint32)
    ids = np.arange(P, dtype=np.uint32)
    for shift, bits in ((0, 11), (11, 11), (22, 10)):
        keys, ids = _stable_pass(keys, ids, shift, bits, blocks=7)
    vis = ids[g["radii"][ids] > 0]
    assert np.array_equal(vis, np.array(sorted(np.where(g["radii"] > 0)[0], key=lambda i: (g["depths"][i].astype(np.float32).view(np.uint32), i)),
                                        dtype=np.uint32))
    # tile pass: instances of a Gaussian = its tile rectangle, row-major (no culling here: the oracle lists every tile of the rectangle)
    def tiles_of(i):
        px, py, r = g["means2D"][i, 0], g["means2D"][i, 1], int(g["radii"][i])
        x0 = min(gx, max(0, int((px - r) / 16))); x1 = min(gx, max(0, int((px + r + 15) / 16)))
        y0 = min(gy, max(0, int((py - r) / 16))); y1 = min(gy, max(0, int((py + r + 15) / 16)))
        return [ty * gx + tx for ty in range(y0, y1) for tx in range(x0, x1)]
    blocks, ntiles = 5, gx * gy
    per = -(-len(ids) // blocks)
    M = np.zeros((blocks, ntiles), np.int64)
    inst = [[(t, int(i)) for i in ids[b * per:(b + 1) * per] if g["radii"][i] > 0 for t in tiles_of(i)] for b in range(blocks)]
    for b in range(blocks):
        for t, _ in inst[b]:
            M[b, t] += 1
    excl = np.cumsum(M.T.reshape(-1)) - M.T.reshape(-1)
    start = excl.reshape(ntiles, blocks).T.copy()
    got = np.zeros(int(M.sum()), np.uint32)
    for b in range(blocks):
        for t, i in inst[b]:
            got[start[b, t]] = i
            start[b, t] += 1
    totals = M.sum(0)
    base = np.cumsum(totals) - totals
    got_ranges = np.where(totals[:, None] > 0, np.stack([base, base + totals], 1), 0).astype(np.uint32)
    assert np.array_equal(got, want_list)
    assert np.array_equal(got_ranges, want_ranges)


def test_single_scalar_recurrence_equals_the_three_colour_recurrences():
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 40))
        alpha = rng.uniform(0.004, 0.99, n)
        col = rng.uniform(0, 1, (n, 3))
        dpix, bg = rng.standard_normal(3), rng.uniform(0, 1, 3)
        T_final = float(np.prod(1 - alpha))
        # reference (backward.cu:470-518), back to front
        T, accum, last_alpha, last_col = T_final, np.zeros(3), 0.0, np.zeros(3)
        ref = np.zeros(n)
        for k in range(n - 1, -1, -1):
            T = T / (1 - alpha[k])
            accum = last_alpha * last_col + (1 - last_alpha) * accum
            last_col, last_alpha = col[k], alpha[k]
            ref[k] = ((col[k] - accum) * dpix).sum() * T + (-T_final / (1 - alpha[k])) * (bg * dpix).sum()
        # ours (lgr_blend.cuh): D = dpix . (background + everything blended behind), in absolute units
        T, D = T_final, T_final * (bg * dpix).sum()
        got = np.zeros(n)
        for k in range(n - 1, -1, -1):
            rcp = 1.0 / (1 - alpha[k])
            T = T * rcp
            cd = (col[k] * dpix).sum()
            got[k] = T * cd - rcp * D
            D += alpha[k] * T * cd
        assert np.allclose(got, ref, rtol=1e-10, atol=1e-12)
