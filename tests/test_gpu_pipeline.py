"""LightGaussian's pipeline end to end on OUR stack, small synthetic scene (the stages the reference chains in
run_prune_finetune.sh / vectree: prune.py -> prune_finetune.py -> vectree.py):
   Global Significance -> volume-weighted score -> prune 50 % (fused compaction, optimizer state carried over)
   -> fine-tune with the fused L1+DSSIM loss and FusedAdamW (quality recovers)
   -> VecTree: importance split, weighted EMA k-means, extreme_saving files -> dequantize -> render (small PSNR cost).
Every stage runs through the rows of SURVEY.md section 8 (a, N1, N2, N3, N4) together."""
import copy
import math

import numpy as np
import pytest
import torch

from lightgaussian_b200 import parallel, optim, vectree
from lightgaussian_b200.loss import l1_ssim_loss
from lightgaussian_b200.model import GaussianParams, TorchCamera, pipeline_params
from lightgaussian_b200.renderer import render, count_render
from lightgaussian_b200.synth import make_scene, make_cameras

pytestmark = pytest.mark.gpu
W, H = 192, 144
NAMES = [("xyz", "_xyz", 1.6e-5), ("f_dc", "_features_dc", 2.5e-3), ("f_rest", "_features_rest", 2.5e-3 / 20), ("opacity", "_opacity", 0.05),
         ("scaling", "_scaling", 0.005), ("rotation", "_rotation", 0.001)]


def psnr(a, b):
    mse = float(((a - b) ** 2).mean())
    return 99.0 if mse == 0 else 10.0 * math.log10(1.0 / mse)


def mean_psnr(pc, cams, targets, pipe, bg):
    with torch.no_grad():
        return float(np.mean([psnr(render(c, pc, pipe, bg)["render"], t) for c, t in zip(cams, targets)]))


def test_prune_finetune_vq_pipeline(tmp_path):
    scene = make_scene(8000, sh_degree=3, seed=51, scale_mult=1.6)
    # real scenes have structured SH (that is what a codebook exploits): draw the colours from a palette of 96 SH vectors + noise
    rng = np.random.default_rng(7)
    palette = np.concatenate([rng.standard_normal((96, 1, 3)), 0.2 * rng.standard_normal((96, 15, 3))], axis=1).astype(np.float32)
    pick = rng.integers(0, 96, 8000)
    sh = palette[pick] + 0.01 * rng.standard_normal((8000, 16, 3)).astype(np.float32)
    scene["raw"]["features_dc"], scene["raw"]["features_rest"] = np.ascontiguousarray(sh[:, :1]), np.ascontiguousarray(sh[:, 1:])
    pc = GaussianParams(scene["raw"], 3, "cuda")
    for _, attr, _ in NAMES:                                   # nn.Parameters, as GaussianModel holds them
        setattr(pc, attr, torch.nn.Parameter(getattr(pc, attr).detach()))
    cams = [TorchCamera(c, "cuda") for c in make_cameras(8, W, H)]
    pipe, bg = pipeline_params(), torch.zeros(3, device="cuda")
    with torch.no_grad():
        targets = [render(c, pc, pipe, bg)["render"].clone() for c in cams]
    pc.optimizer = optim.FusedAdamW([{"params": [getattr(pc, a)], "lr": lr, "name": n} for n, a, lr in NAMES], lr=0.0, eps=1e-15)
    P0 = pc._xyz.shape[0]
    pc.xyz_gradient_accum, pc.denom, pc.max_radii2D = torch.zeros(P0, 1, device="cuda"), torch.zeros(P0, 1, device="cuda"), torch.zeros(P0, device="cuda")

    def iteration(i):
        img = render(cams[i % len(cams)], pc, pipe, bg)["render"]
        loss = l1_ssim_loss(img, targets[i % len(cams)], 0.2)
        loss.backward()
        pc.optimizer.step()
        pc.optimizer.zero_grad(set_to_none=True)
        return float(loss.detach())

    for i in range(4):                                         # a few iterations first: the optimizer has state to carry through the prune
        iteration(i)

    # ---- stage 1: Global Significance pruning (prune.py:112-157, gaussian_model.py:776-782) ----
    with torch.no_grad():
        _, imp = parallel.sharded_prune_list(pc, cams, pipe, bg, count_render, 0, 1)
        volume = torch.prod(pc.get_scaling, dim=1)
        kth = torch.sort(volume, descending=True)[0][int(len(volume) * 0.9)]
        v_score = torch.pow(volume / kth, 0.1) * imp
        thr = torch.sort(v_score)[0][int(0.5 * (P0 - 1))]
        mask = v_score <= thr
    state_before = {n: pc.optimizer.state[getattr(pc, a)]["exp_avg"][~mask].clone() for n, a, _ in NAMES}
    optim.prune_points(pc, mask)
    P1 = pc._xyz.shape[0]
    assert P1 == int((~mask).sum()) and 0.4 * P0 < P1 < 0.6 * P0
    for n, a, _ in NAMES:
        assert getattr(pc, a).shape[0] == P1 and torch.equal(pc.optimizer.state[getattr(pc, a)]["exp_avg"], state_before[n])
    assert pc.denom.shape[0] == P1
    p_pruned = mean_psnr(pc, cams, targets, pipe, bg)

    # ---- stage 2: fine-tune (prune_finetune.py:144-166,287-289) ----
    losses = [iteration(i) for i in range(80)]
    p_tuned = mean_psnr(pc, cams, targets, pipe, bg)
    assert np.mean(losses[-8:]) < 0.8 * np.mean(losses[:8]), losses
    assert p_tuned > p_pruned + 1.0, (p_pruned, p_tuned)

    # ---- stage 3: VecTree (vectree/vectree.py) on the fine-tuned model, importance from a fresh significance pass ----
    with torch.no_grad():
        _, imp = parallel.sharded_prune_list(pc, cams, pipe, bg, count_render, 0, 1)
        feats = torch.cat([pc._xyz, torch.zeros(P1, 3, device="cuda"), pc._features_dc.transpose(1, 2).reshape(P1, -1),
                           pc._features_rest.transpose(1, 2).reshape(P1, -1), pc._opacity, pc._scaling, pc._rotation], dim=1)   # save_ply layout
    torch.manual_seed(0)
    q = vectree.Quantization(feats.cpu().numpy(), importance=imp.cpu().numpy(), sh_degree=3, save_path=str(tmp_path), codebook_size=512,
                             iteration_num=60, vq_ratio=0.6, VQ_CHUNK=4096)
    q.quantize()
    deq = q.dequantize()
    assert deq.shape == feats.shape
    assert torch.equal(deq[:, 0:3], feats[:, 0:3])                                                   # xyz stays float32
    assert torch.equal(deq[:, -8:], feats[:, -8:].half().float())                                    # opacity/scale/rot through fp16
    kept = q.non_vq_mask.cuda()
    assert torch.equal(deq[kept, 6:54], feats[kept, 6:54].half().float())                            # the important 40 % keep their SH (fp16)
    assert not torch.equal(deq[~kept, 6:54], feats[~kept, 6:54].half().float())
    vq = copy.copy(pc)                                                                               # load_ply layout back into the leaves
    vq._xyz = deq[:, 0:3].contiguous()
    vq._features_dc = deq[:, 6:9].reshape(P1, 3, 1).transpose(1, 2).contiguous()
    vq._features_rest = deq[:, 9:54].reshape(P1, 3, 15).transpose(1, 2).contiguous()
    vq._opacity, vq._scaling, vq._rotation = deq[:, 54:55].contiguous(), deq[:, 55:58].contiguous(), deq[:, 58:62].contiguous()
    p_vq = mean_psnr(vq, cams, targets, pipe, bg)
    zero = copy.copy(vq)                                                                             # what losing those SH entirely would cost
    zr = vq._features_rest.clone()
    zr[~kept] = 0
    zero._features_rest = zr
    p_zero = mean_psnr(zero, cams, targets, pipe, bg)
    print(f"PSNR pruned {p_pruned:.2f} tuned {p_tuned:.2f} vq {p_vq:.2f} (SH of the vq'd 60 % zeroed instead: {p_zero:.2f})")
    assert p_vq > p_tuned - 4.0, (p_tuned, p_vq)          # measured: 36.85 -> 34.55 dB with 512 codes / 60 iterations
    assert p_vq > p_zero + 1.0, (p_vq, p_zero)
