"""One training view of the hot path, exactly as the reference's loops drive it
(prune_finetune.py:144-166, distill_train.py:124-166): render() -> image loss -> backward to the raw leaves.
Used by bench.py, smoke() and the tests; the optimizer and SSIM are 'next' rows."""
from __future__ import annotations

import torch


def l1_loss(img: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    return (img - gt).abs().mean()  # utils/loss_utils.py:18-19


def train_view(render_fn, camera, gaussians, pipe, background, target: torch.Tensor):
    """forward + loss + backward for one camera; gradients accumulate into the leaves' .grad.  Returns the loss tensor."""
    pkg = render_fn(camera, gaussians, pipe, background)
    loss = l1_loss(pkg["render"], target)
    loss.backward()
    return loss
