"""Fused image loss of the training loops: the reference's
    loss = (1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))
(prune_finetune.py:160-164, distill_train.py:142-145, utils/loss_utils.py:18-85) in one forward and one backward kernel
(lgr_image_loss_forward / lgr_image_loss_backward) instead of 5 depthwise conv2d + ~15 elementwise kernels and their autograd.
`l1_loss` and `ssim` keep the reference's names and call signatures so that `dropin/utils/loss_utils.py` can export them.
Inputs are float32 CUDA images [3,H,W] (or [C,H,W]); gradients flow to the first argument only (the ground truth /
teacher image never requires grad in the reference's loops)."""
from __future__ import annotations

import torch

from . import capi


def _check(img, gt):
    if not (img.is_cuda and gt.is_cuda and img.dtype == torch.float32 and gt.dtype == torch.float32):
        raise RuntimeError("fused image loss needs float32 CUDA tensors: there is no CPU path")
    if img.dim() != 3 or img.shape != gt.shape:
        raise RuntimeError(f"fused image loss expects two [C,H,W] images of the same shape, got {tuple(img.shape)} and {tuple(gt.shape)}")


def _forward(img, gt, want_maps, l1_only=False):
    lib = capi.load()
    C_, H, W = img.shape
    out = torch.empty(2, dtype=torch.float32, device=img.device)
    ws = torch.empty(int(lib.lgr_image_loss_workspace_bytes(C_, H, W)), dtype=torch.uint8, device=img.device)
    if l1_only:
        if (img.data_ptr() | gt.data_ptr()) & 15:                    # offset views: the vectorised kernel wants 16-byte alignment
            img, gt = img.clone(), gt.clone()
        with torch.cuda.device(img.device):
            st = lib.lgr_image_l1_forward(img.data_ptr(), gt.data_ptr(), C_, H, W, out.data_ptr(), ws.data_ptr(), capi.current_stream_ptr(img.device))
        capi.check(st, "lgr_image_l1_forward")
        return out, None
    dmaps = torch.empty((3, C_, H, W), dtype=torch.float32, device=img.device) if want_maps else None
    with torch.cuda.device(img.device):
        st = lib.lgr_image_loss_forward(img.data_ptr(), gt.data_ptr(), C_, H, W, out.data_ptr(), capi.ptr(dmaps), ws.data_ptr(),
                                        capi.current_stream_ptr(img.device))
    capi.check(st, "lgr_image_loss_forward")
    return out, dmaps


class _ImageLoss(torch.autograd.Function):
    """value = c_l1 * l1 + c_ssim * ssim + c0"""

    @staticmethod
    def forward(ctx, img, gt, c_l1, c_ssim, c0):
        img_c, gt_c = img.contiguous(), gt.contiguous()
        need = ctx.needs_input_grad[0] and c_ssim != 0.0
        out, dmaps = _forward(img_c, gt_c, need, l1_only=(c_ssim == 0.0))
        ctx.save_for_backward(img_c, gt_c, dmaps if dmaps is not None else torch.empty(0, device=img.device))
        ctx.coef = (float(c_l1), float(c_ssim))
        return c_l1 * out[0] + c_ssim * out[1] + c0

    @staticmethod
    def backward(ctx, grad_out):
        img, gt, dmaps = ctx.saved_tensors
        c_l1, c_ssim = ctx.coef
        lib = capi.load()
        C_, H, W = img.shape
        d_img = torch.empty_like(img)
        g = grad_out.contiguous().to(torch.float32)
        with torch.cuda.device(img.device):
            st = lib.lgr_image_loss_backward(img.data_ptr(), gt.data_ptr(), capi.ptr(dmaps), C_, H, W, c_l1, c_ssim, g.data_ptr(),
                                             d_img.data_ptr(), capi.current_stream_ptr(img.device))
        capi.check(st, "lgr_image_loss_backward")
        return d_img, None, None, None, None


def l1_loss(network_output, gt):
    """utils/loss_utils.py:18-19"""
    _check(network_output, gt)
    return _ImageLoss.apply(network_output, gt, 1.0, 0.0, 0.0)


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_utils.py:45-55 (window_size 11, size_average=True: the only form the training loops use)"""
    if window_size != 11 or not size_average:
        raise NotImplementedError("fused ssim implements the training loops' call: window_size=11, size_average=True")
    _check(img1, img2)
    return _ImageLoss.apply(img1, img2, 0.0, 1.0, 0.0)


def l1_ssim_loss(image, gt, lambda_dssim=0.2):
    """(1 - lambda) * L1 + lambda * (1 - SSIM) in ONE forward and ONE backward kernel (prune_finetune.py:160-164)."""
    _check(image, gt)
    return _ImageLoss.apply(image, gt, 1.0 - lambda_dssim, -lambda_dssim, lambda_dssim)
