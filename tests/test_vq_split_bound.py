"""CPU check of the exactness argument behind the tensor-core nearest-code search (csrc/lgr_vq_tc.cuh): with two-term bf16 operands and the
products x_hi.e_hi + x_hi.e_lo + x_lo.e_hi the score error stays far below the decision margin 4e-4 |x| max|e|, so a row that the kernel
decides (second best - best > margin) has the exact float64 argmin.  The bf16 rounding is restated here bit for bit (round to nearest even
on the upper 16 bits, as vt_bf16_rn does)."""
import numpy as np


def bf16_rn(a: np.ndarray) -> np.ndarray:
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32) << 16
    return r.astype(np.uint32).view(np.float32)


def split(a):
    hi = bf16_rn(a)
    lo = bf16_rn((a.astype(np.float32) - hi).astype(np.float32))
    return hi, lo


def coarse_scores(x, e):
    xh, xl = split(x)
    eh, el = split(e)
    # FP32 accumulation of the three partial GEMMs (numpy accumulates in float32 here: at least as coarse as the tensor core's accumulator)
    dot = (xh @ eh.T + xh @ el.T + xl @ eh.T).astype(np.float32)
    nrm = (e.astype(np.float32) ** 2).sum(1, dtype=np.float32)
    return nrm[None, :] - 2.0 * dot


def test_split_error_is_far_below_the_decision_margin():
    rng = np.random.default_rng(0)
    for scale_x, scale_e, d in [(0.5, 0.7, 27), (1.0, 1.5, 27), (3.0, 0.1, 32), (0.05, 5.0, 8)]:
        x = (rng.standard_normal((512, d)) * scale_x).astype(np.float32)
        e = (rng.standard_normal((2048, d)) * scale_e).astype(np.float32)
        approx = coarse_scores(x, e).astype(np.float64)
        x64, e64 = x.astype(np.float64), e.astype(np.float64)
        exact = (e64 ** 2).sum(1)[None, :] - 2.0 * x64 @ e64.T
        xn = np.sqrt((x64 ** 2).sum(1))
        emax = np.sqrt((e64 ** 2).sum(1)).max()
        err = np.abs(approx - exact).max(axis=1)
        margin = 2.0 * 2.0e-4 * xn * emax                 # lgr_vq_tc.cuh: second - best > 2 * VT_MARGIN * |x| * max|e|
        assert np.all(err <= 0.25 * margin), (err / margin).max()   # both scores of a pair may be off: 2 * err <= margin / 2


def test_decided_rows_have_the_exact_argmin():
    rng = np.random.default_rng(1)
    centers = rng.standard_normal((4096, 27)).astype(np.float32)
    x = np.concatenate([(rng.standard_normal((3000, 27)) * 0.5).astype(np.float32),
                        centers[rng.integers(0, 4096, 3000)] + 0.05 * rng.standard_normal((3000, 27)).astype(np.float32),
                        0.5 * (centers[:200] + centers[1:201])])          # samples half way between two codes: near ties by construction
    approx = coarse_scores(x, centers)
    order = np.argsort(approx, axis=1, kind="stable")[:, :2]
    best, second = approx[np.arange(len(x)), order[:, 0]], approx[np.arange(len(x)), order[:, 1]]
    xn = np.sqrt((x.astype(np.float64) ** 2).sum(1))
    emax = np.sqrt((centers.astype(np.float64) ** 2).sum(1)).max()
    decided = (second - best) > 2.0 * 2.0e-4 * xn * emax
    exact = ((centers.astype(np.float64) ** 2).sum(1)[None, :] - 2.0 * x.astype(np.float64) @ centers.astype(np.float64).T).argmin(1)
    assert decided.mean() > 0.5                                        # the coarse pass settles most rows ...
    assert np.array_equal(order[decided, 0], exact[decided])          # ... and every row it settles is settled correctly
    assert not decided[-200:].all()                                    # constructed near ties are handed to the exact kernel
