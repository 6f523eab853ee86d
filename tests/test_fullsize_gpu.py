"""Parity at BASELINE.json's full sizes (192x640 Lite-Mono 3 levels / Monodepth2 4 levels, 384x768 nuScenes shape).

* the fused loss kernels against the CPU oracle on a 2-image batch of the full resolution (the oracle finishes in
  seconds at that size) — loss terms 1e-4 relative (north_star), gradients through oracle/compare.py;
* at the full batch (32 / 16 images) through size-independent properties: the per-level sums are additive over any
  split of the batch and invariant under a permutation of its images (the kernels reduce per image tile, then in a
  fixed order in double precision), the pose gradient of an image does not depend on its batch neighbours;
* the decoder convolutions at the bench's layer shapes against torch's float64 convolution on the same device.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import synth
from oracle import view_synthesis as vs
from oracle.compare import assert_close_robust, check_rel
from oracle import parity_log

pytestmark = pytest.mark.gpu

PHOTO_ONLY = dict(g_d_smooth=0.0, g_c_smooth=0.0, g_m_sparsity=0.0, g_m_smooth=0.0, g_d_ground=0.0)
FRAMES = [-1, 1]

CONFIGS = {   # name: (H, W, scales, phase, intrinsics, ts_mode, full batch)
    "litemono_192x640_fine_tune": (192, 640, [0, 1, 2], "fine_tune", "waymo", "ones", 32),
    "monodepthv2_192x640_disp_init": (192, 640, [0, 1, 2, 3], "disp_init", "kitti", "ones", 16),
    "litemono_384x768_fine_tune": (384, 768, [0, 1, 2], "fine_tune", "nuscenes", "float", 16),
}


class SynthCase:
    def __init__(self, name, batch, seed=11):
        self.H, self.W, self.scales, self.phase, kind, ts_mode, _ = CONFIGS[name]
        self.B = batch
        self.inputs, self.leaves = synth.make_loss_inputs(seed, batch, self.H, self.W, self.scales, kind=kind,
                                                          flow=self.phase != "disp_init", ts_mode=ts_mode)
        synth.add_color_pyramid(self.inputs, self.scales, self.H, self.W)
        self.noise = synth.automask_noise(seed, batch, self.H, self.W, self.scales) if self.phase == "disp_init" else None
        self.step, self.steps_per_epoch = 100, 100

    def tensors(self, device, sel=None, requires_grad=True):
        idx = slice(None) if sel is None else sel
        inputs = {k: (v[idx].to(device).float() if v.is_floating_point() else v[idx].to(device)) for k, v in self.inputs.items()}
        outputs, leaves = {}, {}
        for k, v in self.leaves.items():
            t = v[idx].to(device).float().clone().requires_grad_(requires_grad)
            leaves[k] = t
            if k[0] == "motion_prob":
                for f in FRAMES:
                    outputs[("motion_prob", f, k[1])] = t
                    outputs[("motion_mask", f, k[1])] = torch.sigmoid(t)
            else:
                outputs[k] = t
        noise = None if self.noise is None else {s: n[idx].to(device) for s, n in self.noise.items()}
        return inputs, outputs, leaves, noise


def kernel_sums(case, cfg, inputs, outputs, noise):
    from dd_b200 import functional as Fn

    wc = Fn.WarpConfig(scales=case.scales, cmpflow=cfg.bool_CmpFlow, motmask=cfg.bool_MotMask, automask=cfg.automask)
    disps = [outputs[("disp", 0, s)] for s in case.scales]
    flows = [[outputs[("complete_flow", f, s)] for f in FRAMES] for s in case.scales] if cfg.bool_CmpFlow else None
    masks = [[outputs[("motion_mask", f, s)] for f in FRAMES] for s in case.scales] if cfg.bool_MotMask else None
    noises = [noise[s] for s in case.scales] if noise is not None else None
    return Fn.view_synthesis_sums(wc, inputs[("color", 0, 0)], [inputs[("color", f, 0)] for f in FRAMES], inputs[("K", 0)],
                                  inputs[("inv_K", 0)], [outputs[("cam_T_cam", 0, f)] for f in FRAMES],
                                  [inputs[("ts", f)] for f in FRAMES], disps, flows, masks, noises)


def loss_from_sums(case, cfg, sums, B):
    from dd_b200 import _lib as L

    coef = cfg.coefficients(case.step, case.steps_per_epoch)
    terms = {"p_photo": 0, "c_consistency": 0}
    loss = 0
    for i, s in enumerate(case.scales):
        h, w = case.H >> s, case.W >> s
        photo = sums[i, L.DD_SUM_PHOTO] / (B * case.H * case.W)
        cc = (sums[i, L.DD_SUM_CONSIST0] + sums[i, L.DD_SUM_CONSIST0 + 1]) / (B * 3 * h * w) / (2**s) / 2
        terms["p_photo"] = terms["p_photo"] + photo
        terms["c_consistency"] = terms["c_consistency"] + cc
        lvl = photo * coef["p_photo"] + (cc * coef["c_consistency"] if cfg.bool_MotMask else 0)
        loss = loss + lvl / len(case.scales)
    return loss, terms


def _oracle(case, cfg, dtype):
    o_in, o_out, o_leaves, o_noise = case.tensors("cpu")
    if dtype == torch.float64:
        o_in = {k: (v.double() if v.is_floating_point() else v) for k, v in o_in.items()}
        o_leaves = {k: v.detach().double().requires_grad_(True) for k, v in o_leaves.items()}
        o_out = {}
        for k, t in o_leaves.items():
            if k[0] == "motion_prob":
                for f in FRAMES:
                    o_out[("motion_prob", f, k[1])] = t
                    o_out[("motion_mask", f, k[1])] = torch.sigmoid(t)
            else:
                o_out[k] = t
        o_noise = None if o_noise is None else {s: n.double() for s, n in o_noise.items()}
    vs.generate_images_pred(cfg, o_in, o_out)
    losses = vs.compute_losses(cfg, o_in, o_out, case.step, case.steps_per_epoch, noise=o_noise)
    losses["loss"].backward()
    return losses, {k: v.grad for k, v in o_leaves.items() if v.grad is not None}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_full_resolution_loss_and_grads_vs_oracle(name):
    """Loss terms: 1e-4 relative against the fp32 oracle (north_star).  Gradients: the per-pixel argmin / floor decisions
    make them sensitive to last-ulp differences at this image size (the oracle evaluated in fp32 and in fp64 differs by
    up to ~0.5 % in the pose gradient of the automask configuration), so the bar is "as close to the fp64 oracle as the
    reference's own fp32 arithmetic is": error vs fp64 <= 2x the fp32 oracle's error vs fp64 (floors from oracle/compare.py)."""
    from oracle.compare import robust_report

    case = SynthCase(name, batch=2)
    cfg = vs.LossConfig(case.H, case.W, case.scales, phase=case.phase, **PHOTO_ONLY)
    o_losses, g32 = _oracle(case, cfg, torch.float32)
    _, g64 = _oracle(case, cfg, torch.float64)
    # CUDA path through the C ABI
    g_in, g_out, g_leaves, g_noise = case.tensors("cuda")
    sums = kernel_sums(case, cfg, g_in, g_out, g_noise)
    loss, terms = loss_from_sums(case, cfg, sums, case.B)
    loss.backward()
    torch.cuda.synchronize()
    check_rel(float(loss.detach()), float(o_losses["loss"].detach()), 1e-4, what="loss vs oracle")
    check_rel(float(terms["p_photo"].detach()), float(o_losses["loss_term/p_photo"].detach()), 1e-4, what="p_photo vs oracle")
    if cfg.bool_MotMask:
        check_rel(float(terms["c_consistency"].detach()), float(o_losses["loss_term/c_consistency"].detach()), 1e-4, what="c_consistency vs oracle")
    for k, ref in g64.items():
        got = g_leaves[k].grad
        assert got is not None, k
        pose = k[0] == "cam_T_cam"
        rtol = 2e-3 if pose else 2e-4
        ours, base = robust_report(got.cpu(), ref, rtol), robust_report(g32[k], ref, rtol)
        s_lvl = 0 if pose else k[-1]
        parity_log.record(name, f"grad {k} vs fp64 oracle", rel_l2=ours["rel_l2"], outlier_frac=ours["outlier_frac"],
                          max_abs_rel=ours["max_err"] / ours["scale"], fp32_oracle_rel_l2=base["rel_l2"],
                          fp32_oracle_outlier_frac=base["outlier_frac"])
        # floors: integrated error as in the golden tests; the share of perturbed samples of a level-s map grows like 4^s
        # at a fixed image size (one flipped full-resolution pixel touches ~16 samples of a coarse map)
        assert ours["rel_l2"] <= max(2 * base["rel_l2"], 2e-3 if pose else 2e-2), (k, ours, base)
        if not pose:
            assert ours["outlier_frac"] <= max(2 * base["outlier_frac"], min(2e-3 * 4**s_lvl, 3e-2)), (k, ours, base)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_full_batch_additivity_and_permutation(name):
    full_b = CONFIGS[name][6]
    case = SynthCase(name, batch=full_b, seed=23)
    cfg = vs.LossConfig(case.H, case.W, case.scales, phase=case.phase, **PHOTO_ONLY)
    n_sums = 6

    def run(sel):
        inputs, outputs, leaves, noise = case.tensors("cuda", sel)
        sums = kernel_sums(case, cfg, inputs, outputs, noise)
        b = inputs[("color", 0, 0)].shape[0]
        loss, _ = loss_from_sums(case, cfg, sums, b)
        (loss * b).backward()     # un-normalised so that per-image gradients are comparable across batch sizes
        return sums.detach()[:, :n_sums].double().cpu(), {k: v.grad.detach().cpu() for k, v in leaves.items() if v.grad is not None}

    s_all, g_all = run(None)
    half = full_b // 2
    s_a, g_a = run(slice(0, half))
    s_b, g_b = run(slice(half, full_b))
    scale = s_all.abs().clamp_min(1e-12)
    assert ((s_a + s_b - s_all).abs() / scale).max().item() <= 1e-5
    perm = torch.randperm(full_b, generator=torch.Generator().manual_seed(5))
    s_p, g_p = run(perm)
    assert ((s_p - s_all).abs() / scale).max().item() <= 1e-5
    # per-image gradients do not depend on the batch an image sits in
    for k, g in g_all.items():
        assert_close_robust(torch.cat([g_a[k], g_b[k]]), g, rtol=1e-5, max_outlier_frac=1e-4, max_rel_l2=1e-5, what=("split", k))
        assert_close_robust(g_p[k], g[perm], rtol=1e-5, max_outlier_frac=1e-4, max_rel_l2=1e-5, what=("perm", k))


BENCH_LAYERS = [   # C0, C1, Cout, H, W (output), k, pad, act, up, batch
    (112, 128, 112, 24, 80, 3, "reflect", "elu", "bilinear", 32),     # Lite decoder .1
    (64, 64, 64, 48, 160, 3, "reflect", "elu", "bilinear", 32),       # Lite decoder .3
    (3, 64, 64, 96, 320, 3, "zero", "none", "none", 32),              # motion level 4 conv0
    (64, 0, 64, 96, 320, 3, "zero", "none", "none", 32),              # motion level 4 conv1
    (3, 9, 9, 192, 640, 3, "zero", "none", "none", 32),               # motion level 5 conv0
    (64, 64, 3, 96, 320, 1, "zero", "none", "none", 32),              # motion level 4 reduction
    (32, 0, 1, 96, 320, 3, "reflect", "none", "none", 32),            # dispconv 0
    (256, 0, 256, 12, 40, 3, "reflect", "elu", "nearest", 16),        # Monodepth2 upconv_4_1 (first half)
    (64, 64, 64, 192, 384, 3, "zero", "none", "none", 8),             # motion level 4 at 384x768
]


def _torch_conv(x0, x1, w, b, ks, pad, act, up):
    x = x0
    if up == "nearest":
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif up == "bilinear":
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    if x1 is not None:
        x = torch.cat([x, x1], 1)
    if ks == 3:
        x = F.pad(x, (1, 1, 1, 1), mode="reflect" if pad == "reflect" else "constant")
    y = F.conv2d(x, w, b)
    return F.elu(y) if act == "elu" else y


@pytest.mark.parametrize("layer", BENCH_LAYERS, ids=lambda l: "x".join(str(v) for v in l))
def test_bench_layers_vs_torch_fp64(layer):
    from dd_b200 import functional as Fn

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    C0, C1, Cout, h, w, ks, pad, act, up, B = layer
    g = torch.Generator(device="cuda").manual_seed(C0 * 7 + Cout)
    h0, w0 = (h, w) if up == "none" else (h // 2, w // 2)
    x0 = torch.randn(B, C0, h0, w0, device="cuda", generator=g)
    x1 = torch.randn(B, C1, h, w, device="cuda", generator=g) if C1 else None
    wt = torch.randn(Cout, C0 + C1, ks, ks, device="cuda", generator=g) * (1.0 / (ks * (C0 + C1) ** 0.5))
    bias = torch.randn(Cout, device="cuda", generator=g) * 0.1
    go = None
    res = []
    for impl in ("ours", "torch64"):
        dt = torch.float32 if impl == "ours" else torch.float64
        leaves = [t.to(dt).clone().requires_grad_(True) if t is not None else None for t in (x0, x1, wt, bias)]
        a0, a1, aw, ab = leaves
        out = Fn.conv2d_fused(a0, aw, ab, x1=a1, ksize=ks, pad=pad, act=act, up=up) if impl == "ours" else \
            _torch_conv(a0, a1, aw, ab, ks, pad, act, up)
        if go is None:
            go = torch.randn(out.shape, device="cuda", generator=g)
        out.backward(go.to(dt))
        res.append([out.detach()] + [t.grad for t in leaves if t is not None])
    torch.cuda.synchronize()
    names = ["out", "grad_x0"] + (["grad_x1"] if C1 else []) + ["grad_w", "grad_b"]
    for n, a, b_ in zip(names, res[0], res[1]):
        a = a.double()
        rel = ((a - b_).norm() / b_.norm().clamp_min(1e-30)).item()
        parity_log.record("bench_layer_" + "x".join(str(v) for v in layer), n, rel_l2=rel,
                          max_abs_rel=((a - b_).abs().max() / b_.abs().max().clamp_min(1e-30)).item())
        # against the float64 result of the same op: fp32 rounding only (Winograd transforms included); the weight and
        # bias gradients sum ~1e6 products per element in fp32
        assert rel <= (3e-5 if n in ("grad_w", "grad_b") else 1e-5), (n, rel)
        assert ((a - b_).abs().max() / b_.abs().max().clamp_min(1e-30)).item() <= 1e-4, n
