"""The reference's scripts, UNMODIFIED, executed on the GPU through the drop-ins (north_star: "prune_finetune.py, distill_train.py and
render_video.py run unmodified") and, side by side, through the reference's own stock stack (its pybind extension built from
RAST/setup.py, its gaussian_renderer, its torch loss and torch.optim.AdamW).

Reference loops exercised: prune_finetune.py:99-289 (render -> L1 + DSSIM -> backward -> AdamW, prune at the first iteration through
prune.prune_list / calculate_v_imp_score, imp_score.npz at the last checkpoint), distill_train.py:100-183 with the C4 flags
(--new_max_sh 2 --augmented_view --enable_covariance: the student's _features_rest is the NON-CONTIGUOUS [P,8,3] view of
scene/gaussian_model.py:129-136), render_video.py:107-158, prune.py:133-157.

Pass criteria: same first-iteration loss to 1e-6 (absolute; the losses are O(0.05)), loss curve and final PSNR within 3 %, significance
counts equal to the exact oracle, identical rendered frames.
"""
import os

import numpy as np
import pytest

from tests import scripts_harness as sh

pytestmark = pytest.mark.gpu

ITER0 = 30000
STEPS = 200          # iterations each training script runs
PORT = 6100


@pytest.fixture(scope="module")
def work(tmp_path_factory):
    reason = sh.stacks_available()
    if reason:
        pytest.skip(reason)
    base = str(tmp_path_factory.mktemp("lgscripts"))
    w = sh.build_workdir(base, iteration=ITER0)
    w["gt_q"] = [np.rint(np.clip(g, 0, 1) * 255.0).astype(np.float32) / 255.0 for g in w["gt"]]
    w["runs"] = {}
    return w


def _curve(model_dir):
    c = sh.read_scalars(model_dir, "train_loss_patches/total_loss")
    steps = sorted(c)
    return steps, np.array([c[s] for s in steps])


def _prune_finetune(work, stack):
    key = ("prune_finetune", stack)
    if key in work["runs"]:
        return work["runs"][key]
    out = os.path.join(work["base"], f"pf_{stack}")
    trace = os.path.join(work["base"], f"pf_{stack}.trace.json")
    last = ITER0 + STEPS
    # the C3 flags of scripts/run_prune_finetune.sh:37-48,71-82; --test_iterations unreachable (LPIPS would download VGG weights)
    sh.run(stack, ["prune_finetune.py", "-s", work["data"], "-m", out, "--eval", "-r", "1", "--port", str(PORT + (stack == "ours")),
                   "--start_checkpoint", work["ckpt"], "--iterations", str(last), "--prune_percent", "0.66", "--prune_type", "v_important_score",
                   "--prune_decay", "1", "--v_pow", "0.1", "--position_lr_max_steps", str(last), "--prune_iterations", str(ITER0 + 1),
                   "--test_iterations", "999999", "--save_iterations", str(last), "--checkpoint_iterations", str(last)], trace=trace)
    res = dict(out=out, trace=sh.read_trace(trace) if stack == "ours" else None, last=last)
    work["runs"][key] = res
    return res


def test_prune_finetune_runs_unmodified_and_matches_the_stock_stack(work):
    ours, stock = _prune_finetune(work, "ours"), _prune_finetune(work, "stock")
    so, co = _curve(ours["out"])
    ss, cs = _curve(stock["out"])
    assert so == ss and len(so) == STEPS and so[0] == ITER0 + 1
    # iteration 1 renders the unpruned checkpoint: forward is bit-identical, the fused loss is within float rounding of torch's
    assert abs(co[0] - cs[0]) <= 1e-6, (co[0], cs[0])
    # the prune (iteration 1, after the loss) removes 66 %: the loss jumps, then fine-tuning brings it down again on both stacks
    assert co[1:6].mean() > 1.2 * co[0] and cs[1:6].mean() > 1.2 * cs[0]
    tail_o, tail_s = co[-50:].mean(), cs[-50:].mean()
    assert tail_o < 0.9 * co[1:11].mean() and tail_s < 0.9 * cs[1:11].mean()
    # After the prune the two stacks train DIFFERENT survivor sets: the stock stack ranks by its racy, non-atomic counter
    # (forward.cu:473-474 loses most increments), ours by the exact count.  So the criterion is one-sided:
    # our curve must not be worse than the stock stack's by more than 3 % (smoothed over 20 iterations: single iterations differ
    # with the random camera's difficulty); both values are printed.
    win = lambda c: np.convolve(c[1:], np.ones(20) / 20, mode="valid")  # noqa: E731
    rel = (win(co) - win(cs)) / win(cs)
    print(f"prune_finetune tail loss ours {tail_o:.5f} stock {tail_s:.5f}; smoothed (ours-stock)/stock in [{rel.min():+.3f}, {rel.max():+.3f}]")
    assert tail_o <= 1.03 * tail_s, (tail_o, tail_s)
    assert rel.max() <= 0.03, rel.max()
    # our stack really took the fused kernels for every call, and the fused optimizer
    t = ours["trace"]
    assert t.get("render_unfused", 0) == 0 and t["render_fused"] >= STEPS and t["adamw_steps"] == STEPS - 1, t
    # final PSNR on the held-out views, both results rendered by the same renderer
    test_idx = [k for k in range(len(work["cams"])) if k % 8 == 0]
    cams, gt = [work["cams"][k] for k in test_idx], [work["gt_q"][k] for k in test_idx]
    p = {}
    for name, r in (("ours", ours), ("stock", stock)):
        ck = sh.load_checkpoint_leaves(os.path.join(r["out"], f"chkpnt{r['last']}.pth"))
        assert ck["leaves"]["xyz"].shape[0] == work["P"] - int(0.66 * work["P"]) or abs(ck["leaves"]["xyz"].shape[0] - 0.34 * work["P"]) <= 2
        p[name] = sh.psnr_of_leaves(ck["leaves"], 3, cams, gt)
    print(f"prune_finetune held-out PSNR ours {p['ours']:.3f} dB, stock {p['stock']:.3f} dB")
    assert p["ours"] >= 0.97 * p["stock"], p
    work["runs"]["psnr_pf"] = p


def test_imp_score_written_by_the_script_equals_the_exact_oracle(work):
    """prune_finetune.py:205-210 writes imp_score.npz = calculate_v_imp_score(prune_list(...)) at the last checkpoint.  Recompute it
    from the saved checkpoint with the CPU oracle's exact counts: values equal to float rounding, ranking identical."""
    from oracle.lgo import Oracle
    from tests.util import view_from_camera
    ours = _prune_finetune(work, "ours")
    ck = sh.load_checkpoint_leaves(os.path.join(ours["out"], f"chkpnt{ours['last']}.pth"))
    L = ck["leaves"]
    rot = L["rotation"]
    act = dict(means3D=L["xyz"], scales=np.exp(L["scaling"]), rotations=rot / np.sqrt((rot * rot).sum(1, keepdims=True)),
               opacities=(1.0 / (1.0 + np.exp(-L["opacity"].astype(np.float64)))).astype(np.float32),
               shs=np.ascontiguousarray(np.concatenate([L["features_dc"], L["features_rest"]], axis=1)))
    # the oracle must see the activations the GPU computed: take them from torch on the GPU (bit-identical to the kernels' own)
    import torch
    act["scales"] = torch.exp(torch.from_numpy(L["scaling"]).cuda()).cpu().numpy()
    act["rotations"] = torch.nn.functional.normalize(torch.from_numpy(rot).cuda()).cpu().numpy()
    act["opacities"] = torch.sigmoid(torch.from_numpy(L["opacity"]).cuda()).cpu().numpy()
    train = [c for k, c in enumerate(work["cams"]) if k % 8 != 0]
    o = Oracle()
    counts = np.zeros(L["xyz"].shape[0], np.int64)
    fragile = 0
    for c in train:
        f = o.forward(view_from_camera(c, (0, 0, 0), 3, 1.0), act["means3D"], act["opacities"], shs=act["shs"], scales=act["scales"],
                      rotations=act["rotations"], count=True, want_fragile=True)
        counts += f["gaussians_count"]
        fragile += int(f["fragile"].sum())
    vol = np.prod(act["scales"].astype(np.float64), axis=1)
    kth = np.sort(vol)[::-1][int(len(vol) * 0.9)]
    expect = (vol / kth) ** 0.1 * (act["opacities"][:, 0].astype(np.float64) * counts)
    got = np.load(os.path.join(ours["out"], "imp_score.npz"))["arr_0"].astype(np.float64)
    assert got.shape == expect.shape
    # pixels whose threshold tests sit within exp() rounding (glibc vs MUFU) may flip one pair each: bounded by the fragile count
    bad = np.abs(got - expect) > 2e-5 * np.maximum(expect, 1e-6)
    assert bad.sum() <= max(4, 2 * fragile), (int(bad.sum()), fragile)
    order = np.argsort(got, kind="stable")
    e = expect[order]
    inversions = (e[:-1] > e[1:] * (1 + 1e-5) + 1e-9)
    assert inversions.sum() <= max(4, 2 * fragile), int(inversions.sum())


def test_prune_list_through_the_reference_module(work):
    """prune.py:133-157 driven exactly as prune_finetune.py does, on both stacks: our counts are the exact integers; the stock stack's
    non-atomic `count++` (forward.cu:473) can only lose increments."""
    outs = {}
    for stack in ("ours", "stock"):
        out = os.path.join(work["base"], f"prune_list_{stack}.npz")
        sh.run(stack, [os.path.join(sh.HELPERS, "run_prune_list.py"), "-s", work["data"], "-m", os.path.join(work["base"], f"pl_{stack}"),
                       "--eval", "-r", "1", "--start_checkpoint", work["ckpt"], "--out", out])
        outs[stack] = np.load(out)
    co, cs = outs["ours"]["count"].astype(np.int64), outs["stock"]["count"].astype(np.int64)
    assert co.shape == (work["P"],) and co.sum() > 0
    assert (cs <= co).all() and cs.sum() > 0
    # ours equals the exact oracle count (sum over the train views), integer for integer
    from oracle.lgo import Oracle
    from tests.util import view_from_camera
    act = work["act"]
    o, exact, fragile = Oracle(), np.zeros(work["P"], np.int64), 0
    for c in [c for k, c in enumerate(work["cams"]) if k % 8 != 0]:
        f = o.forward(view_from_camera(c, (0, 0, 0), 3, 1.0), act["means3D"], act["opacities"], shs=act["shs"], scales=act["scales"],
                      rotations=act["rotations"], count=True, want_fragile=True)
        exact += f["gaussians_count"]
        fragile += int(f["fragile"].sum())
    assert np.abs(co - exact).sum() <= 64 * fragile, (int(np.abs(co - exact).sum()), fragile)
    if fragile == 0:
        np.testing.assert_array_equal(co, exact)
    # ranking by the volume-weighted score: how far the stock stack's racy counter moves the prune decision (reported)
    k = int(0.66 * work["P"])
    drop_o, drop_s = set(np.argsort(outs["ours"]["v"])[:k].tolist()), set(np.argsort(outs["stock"]["v"])[:k].tolist())
    print(f"prune decision overlap ours/stock: {len(drop_o & drop_s) / k:.3f}; stock counter keeps {cs.sum() / co.sum():.3f} of the increments")
    assert len(drop_o & drop_s) >= 0.5 * k


def test_distill_train_runs_unmodified_with_the_strided_student(work):
    runs = {}
    last = ITER0 + STEPS
    for stack in ("ours", "stock"):
        out = os.path.join(work["base"], f"distill_{stack}")
        trace = os.path.join(work["base"], f"distill_{stack}.trace.json")
        # the C4 flags of scripts/run_distill_finetune.sh:41-52
        sh.run(stack, ["distill_train.py", "-s", work["data"], "-m", out, "--eval", "-r", "1", "--port", str(PORT + 10 + (stack == "ours")),
                       "--start_checkpoint", work["ckpt"], "--teacher_model", work["ckpt"], "--iterations", str(last), "--new_max_sh", "2",
                       "--position_lr_max_steps", str(last), "--enable_covariance", "--augmented_view",
                       "--test_iterations", "999999", "--save_iterations", str(last), "--checkpoint_iterations", str(last)], trace=trace)
        runs[stack] = dict(out=out, trace=sh.read_trace(trace) if stack == "ours" else None)
    so, co = _curve(runs["ours"]["out"])
    ss, cs = _curve(runs["stock"]["out"])
    assert so == ss and len(so) == STEPS
    # same renders bit for bit; the fused L1 + SSIM differs from torch's conv2d composition by float rounding only (both ~1e-6 off float64)
    assert abs(co[0] - cs[0]) <= 1e-6, (co[0], cs[0])
    tail_o, tail_s = co[-50:].mean(), cs[-50:].mean()
    assert tail_o < co[:10].mean() and tail_s < cs[:10].mean()          # distillation converges towards the teacher
    assert abs(tail_o - tail_s) <= 0.03 * tail_s, (tail_o, tail_s)
    t = runs["ours"]["trace"]
    # student (strided M=9 leaf) and teacher both went through the fused kernels; the optimizer updated the strided leaf in place
    assert t.get("render_unfused", 0) == 0 and t["render_fused"] >= 2 * STEPS and t["render_fused_strided_rest"] >= STEPS, t
    assert t["adamw_steps"] == STEPS - 1 and t["adamw_strided_params"] == STEPS - 1, t
    cko = sh.load_checkpoint_leaves(os.path.join(runs["ours"]["out"], f"chkpnt{last}.pth"))
    cks = sh.load_checkpoint_leaves(os.path.join(runs["stock"]["out"], f"chkpnt{last}.pth"))
    assert cko["leaves"]["features_rest"].shape == (work["P"], 8, 3) == cks["leaves"]["features_rest"].shape
    assert cko["active_sh_degree"] == 2
    cams, gt = work["cams"][::8], work["gt_q"][::8]
    po, ps = sh.psnr_of_leaves(cko["leaves"], 2, cams, gt), sh.psnr_of_leaves(cks["leaves"], 2, cams, gt)
    assert abs(po - ps) <= 0.03 * ps, (po, ps)
    # imp_score.npz of the student (distill_train.py:166-176) exists on both
    for s in ("ours", "stock"):
        assert np.load(os.path.join(runs[s]["out"], "imp_score.npz"))["arr_0"].shape == (work["P"],)


def test_render_video_runs_unmodified_and_frames_are_identical(work):
    """render_video.py on the model prune_finetune.py saved (point_cloud/iteration_N/point_cloud.ply + cfg_args), both stacks on the
    SAME model directory contents: the forward pass is bit-identical, so the PNG frames are equal."""
    import shutil
    from PIL import Image
    src = _prune_finetune(work, "ours")
    frames = {}
    for stack in ("ours", "stock"):
        model = os.path.join(work["base"], f"rv_{stack}")
        shutil.copytree(src["out"], model)
        # cfg_args stores model_path: point it at the copy
        cfg = open(os.path.join(model, "cfg_args")).read().replace(src["out"], model)
        open(os.path.join(model, "cfg_args"), "w").write(cfg)
        sh.run(stack, ["render_video.py", "-m", model, "--skip_test", "--quiet"])
        rdir = os.path.join(model, "train", f"ours_{src['last']}", "renders")
        names = sorted(os.listdir(rdir))
        assert len(names) == len(work["cams"]) - len(work["cams"][::8])
        frames[stack] = [np.asarray(Image.open(os.path.join(rdir, n))) for n in names]
    for a, b in zip(frames["ours"], frames["stock"]):
        assert a.shape == b.shape == (work["H"], work["W"], 3)
        assert np.array_equal(a, b)
