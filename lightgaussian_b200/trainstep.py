"""One training view of the hot path, exactly as the reference's loops drive it
(prune_finetune.py:144-166, distill_train.py:124-166): render() -> image loss -> backward to the raw leaves.
Used by bench.py, smoke() and the tests."""
from __future__ import annotations

import torch


def l1_loss(img: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    return (img - gt).abs().mean()  # utils/loss_utils.py:18-19


def train_view(render_fn, camera, gaussians, pipe, background, target: torch.Tensor, loss_fn=None):
    """forward + loss + backward for one camera; gradients accumulate into the leaves' .grad.  Returns the loss tensor.
    loss_fn defaults to the reference's torch composition of l1_loss; our stack passes lightgaussian_b200.loss.l1_loss (same value)."""
    pkg = render_fn(camera, gaussians, pipe, background)
    loss = (loss_fn or l1_loss)(pkg["render"], target)
    loss.backward()
    return loss
