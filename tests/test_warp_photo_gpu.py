"""GPU parity of the fused view-synthesis + photometric kernels (through the C ABI / ctypes) against
the CPU oracle and the goldens recorded from the reference.  Tolerance: 1e-4 relative fp32 on
integrated quantities (north_star); per-pixel tensors allow the argmin/floor flips described in
oracle/compare.py."""
import pytest
import torch

from oracle import view_synthesis as vs
from oracle.compare import assert_close_robust, check_rel, grad_bounds
from oracle.golden_io import LOSS_CASE_NAMES, LossCase

pytestmark = pytest.mark.gpu

PHOTO_ONLY = dict(g_d_smooth=0.0, g_c_smooth=0.0, g_m_sparsity=0.0, g_m_smooth=0.0, g_d_ground=0.0)


def oracle_photo(case, dtype=torch.float32):
    cfg = vs.LossConfig(case.H, case.W, case.scales, phase=case.phase, **PHOTO_ONLY)
    inputs = case.cast_inputs(dtype)
    outputs, leaves = case.fresh_outputs(dtype)
    vs.generate_images_pred(cfg, inputs, outputs)
    losses = vs.compute_losses(cfg, inputs, outputs, case.step, case.steps_per_epoch, noise=case.noise)
    losses["loss"].backward()
    return cfg, outputs, leaves, losses


def cuda_photo(case, cfg, materialise=(), keep_warped=True):
    from dd_b200 import functional as Fn
    from dd_b200 import _lib as L

    dev = "cuda"
    inputs = case.cast_inputs(torch.float32, dev)
    outputs, leaves = case.fresh_outputs(torch.float32, dev)
    frames = [-1, 1]
    wc = Fn.WarpConfig(scales=case.scales, cmpflow=cfg.bool_CmpFlow, motmask=cfg.bool_MotMask, automask=cfg.automask,
                       materialise=tuple(materialise), keep_warped=keep_warped)
    disps = [outputs[("disp", 0, s)] for s in case.scales]
    flows = [[outputs[("complete_flow", f, s)] for f in frames] for s in case.scales] if cfg.bool_CmpFlow else None
    masks = [[outputs[("motion_mask", f, s)] for f in frames] for s in case.scales] if cfg.bool_MotMask else None
    noises = [case.noise[s].to(dev) for s in case.scales] if case.noise is not None else None
    sums = Fn.view_synthesis_sums(wc, inputs[("color", 0, 0)], [inputs[("color", f, 0)] for f in frames], inputs[("K", 0)],
                                  inputs[("inv_K", 0)], [outputs[("cam_T_cam", 0, f)] for f in frames],
                                  [inputs[("ts", f)] for f in frames], disps, flows, masks, noises)
    coef = cfg.coefficients(case.step, case.steps_per_epoch)
    B, H, W = case.B, case.H, case.W
    terms = {"p_photo": 0, "c_consistency": 0}
    loss = 0
    for i, s in enumerate(case.scales):
        h, w = H >> s, W >> s
        photo = sums[i, L.DD_SUM_PHOTO] / (B * H * W)
        cc = (sums[i, L.DD_SUM_CONSIST0] + sums[i, L.DD_SUM_CONSIST0 + 1]) / (B * 3 * h * w) / (2**s) / 2
        terms["p_photo"] = terms["p_photo"] + photo
        terms["c_consistency"] = terms["c_consistency"] + cc
        lvl = photo * coef["p_photo"]
        if cfg.bool_MotMask:
            lvl = lvl + cc * coef["c_consistency"]
        loss = loss + lvl / len(case.scales)
    loss.backward()
    return wc, leaves, terms, loss, sums


@pytest.mark.parametrize("keep_warped", [True, False], ids=["saved_warp", "recompute"])
@pytest.mark.parametrize("name", LOSS_CASE_NAMES)
def test_photo_loss_and_grads_vs_oracle(name, keep_warped):
    case = LossCase(name)
    cfg, o_out, o_leaves, o_losses = oracle_photo(case)
    wc, leaves, terms, loss, sums = cuda_photo(case, cfg, keep_warped=keep_warped)
    torch.cuda.synchronize()
    check_rel(float(loss), float(o_losses["loss"]), 1e-4, what="loss vs oracle")
    check_rel(float(terms["p_photo"]), float(o_losses["loss_term/p_photo"]), 1e-4, what="p_photo vs oracle")
    # the golden from the reference holds the same photometric term
    check_rel(float(terms["p_photo"]), case.losses["loss_term/p_photo"], 1e-4, what="p_photo vs reference golden")
    if cfg.bool_MotMask:
        check_rel(float(terms["c_consistency"]), float(o_losses["loss_term/c_consistency"]), 1e-4, what="c_consistency vs oracle")
        check_rel(float(terms["c_consistency"]), case.losses["loss_term/c_consistency"], 1e-4, what="c_consistency vs reference golden")
    for k, ref in o_leaves.items():
        if ref.grad is None:
            continue
        got = leaves[k].grad
        assert got is not None, k
        bounds = grad_bounds(name, k)
        if name == "loss_dispinit_lite_96x128":
            # here the comparison partner is the ORACLE, whose own automask argmin flips against the reference on this case
            # (oracle vs reference golden: 8e-3 rel-L2 on the level-0 map, 8e-4 on the pose gradient; CUDA path vs the same
            # golden: 1.3e-5 / 1.4e-4, tests/test_trainer_losses_gpu.py)
            bounds = dict(rtol=3e-3, max_outlier_frac=0.0, max_rel_l2=2.5e-3) if k[0] == "cam_T_cam" else \
                dict(rtol=1e-4, max_outlier_frac=1.2e-3, max_rel_l2=2.5e-2)
        assert_close_robust(got.cpu(), ref.grad, what=k, **bounds)


@pytest.mark.parametrize("name", [n for n in LOSS_CASE_NAMES if "32x64" in n])
def test_materialised_outputs_vs_reference_golden(name):
    case = LossCase(name)
    cfg = vs.LossConfig(case.H, case.W, case.scales, phase=case.phase, **PHOTO_ONLY)
    wc, leaves, terms, loss, sums = cuda_photo(case, cfg, materialise=("warped", "sample", "depth", "ident_sel", "resid", "independ"))
    torch.cuda.synchronize()
    frames = [-1, 1]
    checked = 0
    for (name_, f, i), t in wc.aux.items():
        s = case.scales[i]
        key = {"warped": ("color", frames[f], s), "sample": ("sample", frames[f], s), "depth": ("depth", 0, s),
               "resid": ("residual_flow", frames[f], s), "independ": ("independ_flow", frames[f], s),
               "ident_sel": f"identity_selection/{s}"}.get(name_)
        if key is None or key not in case.outputs:
            continue
        ref = case.outputs[key]
        got = t.cpu()
        if name_ == "ident_sel":
            assert (got != ref).float().mean().item() < 2e-3, key
        else:
            assert_close_robust(got, ref, rtol=3e-4, max_outlier_frac=2e-3, max_rel_l2=1e-3, what=key)
        checked += 1
    assert checked >= 6


def test_cpu_tensors_are_rejected():
    from dd_b200 import functional as Fn
    from dd_b200 import DynamoB200Error

    case = LossCase("loss_dispinit_md2_32x64")
    inputs = case.cast_inputs(torch.float32, "cpu")
    outputs, _ = case.fresh_outputs(torch.float32, "cpu")
    wc = Fn.WarpConfig(scales=case.scales, automask=True)
    with pytest.raises(DynamoB200Error):
        Fn.view_synthesis_sums(wc, inputs[("color", 0, 0)], [inputs[("color", f, 0)] for f in (-1, 1)], inputs[("K", 0)],
                               inputs[("inv_K", 0)], [outputs[("cam_T_cam", 0, f)] for f in (-1, 1)], [None, None],
                               [outputs[("disp", 0, s)] for s in case.scales])
