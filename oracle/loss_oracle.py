"""CPU oracle of the training image loss (TEST INFRASTRUCTURE ONLY): numpy float64 restatement of the reference's
utils/loss_utils.py:18-85 (l1_loss, ssim with the 11x11 Gaussian window, zero padding) and of the analytic gradient of
    c_l1 * l1 + c_ssim * ssim   with respect to the first image.
Pinned by tests/golden/pyref_loss.npz (outputs and autograd gradients of the reference's own module, imported on CPU by
tests/golden/make_pyref_loss_golden.py) and by a finite-difference check."""
from math import exp

import numpy as np


def window_1d():
    g = np.array([exp(-((x - 5) ** 2) / float(2 * 1.5 ** 2)) for x in range(11)], dtype=np.float32)   # loss_utils.py:26-33
    return (g / g.sum(dtype=np.float32)).astype(np.float32)


def window_2d():
    g = window_1d()
    return (g[:, None] * g[None, :]).astype(np.float32)          # _1D_window.mm(_1D_window.t()), loss_utils.py:36-38


def conv_same(img, w2d):
    """per-channel correlation with zero padding 5 (F.conv2d(..., padding=5, groups=C))"""
    C, H, W = img.shape
    pad = np.zeros((C, H + 10, W + 10), np.float64)
    pad[:, 5:5 + H, 5:5 + W] = img
    out = np.zeros((C, H, W), np.float64)
    for i in range(11):
        for j in range(11):
            out += float(w2d[i, j]) * pad[:, i:i + H, j:j + W]
    return out


def l1_ssim(x, y):
    """returns (l1, ssim, maps) ; maps = (A, B, C) partial-derivative maps used by the gradient"""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    w = window_2d()
    mu1, mu2 = conv_same(x, w), conv_same(y, w)
    s1 = conv_same(x * x, w) - mu1 * mu1
    s2 = conv_same(y * y, w) - mu2 * mu2
    s12 = conv_same(x * y, w) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    a1, a2, b1, b2 = 2 * mu1 * mu2 + C1, 2 * s12 + C2, mu1 * mu1 + mu2 * mu2 + C1, s1 + s2 + C2
    m = a1 * a2 / (b1 * b2)
    dB = -a1 * a2 / (b1 * b2 * b2)
    dC = 2 * a1 / (b1 * b2)
    dmu1 = 2 * mu2 * a2 / (b1 * b2) - 2 * mu1 * a1 * a2 / (b1 * b1 * b2)
    A = dmu1 - 2 * mu1 * dB - mu2 * dC
    return float(np.abs(x - y).mean()), float(m.mean()), (A, dB, dC)


def grad_wrt_first(x, y, c_l1, c_ssim):
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    _, _, (A, B, Cm) = l1_ssim(x, y)
    w = window_2d()
    n = x.size
    g = c_ssim * (conv_same(A, w) + 2 * x * conv_same(B, w) + y * conv_same(Cm, w))
    g = g + c_l1 * np.sign(x - y)
    return g / n
