"""Development micro-benchmark: times the fused warp+SSIM kernels and representative decoder
convolutions at the bench shapes with CUDA events (also the target of `ncu -k regex:...`).

    python dev/kernel_bench.py [--what warp|conv|all] [--reps 10] [--batch 32]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dynamo-depth_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from dd_b200 import functional as Fn  # noqa: E402
from dd_b200 import synthetic  # noqa: E402
import options  # noqa: E402

H, W = 192, 640


def ev_time(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def warp_inputs(B, scales, mode, dev):
    opt = options.DynamoOptions().parse(args=["-d", "waymo", "--depth_model", "litemono", "--weights_init", "scratch", "--height",
                                              str(H), "--width", str(W), "-b", str(B)])
    batch = {k: v.to(dev) for k, v in synthetic.make_batch(opt, 7).items()}
    g = torch.Generator(device=dev).manual_seed(3)
    disps = [(0.002 + 0.098 * torch.rand(B, 1, H >> s, W >> s, device=dev, generator=g)).requires_grad_(True) for s in scales]
    Ts = []
    for sgn in (-1.0, 1.0):
        T = torch.eye(4, device=dev).repeat(B, 1, 1)
        T[:, :3, 3] = 0.1 * torch.randn(B, 3, device=dev, generator=g)
        T[:, 2, 3] += 0.5 * sgn
        Ts.append(T.requires_grad_(True))
    flows = masks = None
    if mode >= 1:
        flows = [[(0.05 * torch.randn(B, 3, H >> s, W >> s, device=dev, generator=g)).requires_grad_(True) for _ in range(2)] for s in scales]
    if mode == 2:
        masks = [[torch.rand(B, 1, H >> s, W >> s, device=dev, generator=g).requires_grad_(True) for _ in range(2)] for s in scales]
    return batch, disps, Ts, flows, masks


def bench_warp(B, reps, only=None):
    dev = torch.device("cuda")
    scales = [0, 1, 2]
    P = H * W
    for name, mode, automask in (("disp_init(rigid+automask)", 0, True), ("motion_init(flow)", 1, False), ("fine_tune(flow+mask)", 2, False)):
        if only is not None and mode != only:
            continue
        batch, disps, Ts, flows, masks = warp_inputs(B, scales, mode, dev)
        cfg = Fn.WarpConfig(scales=scales, cmpflow=mode >= 1, motmask=mode == 2, automask=automask,
                            keep_warped=os.environ.get("DD_BENCH_RECOMPUTE") is None)
        noises = [torch.randn(B, 2, H, W, device=dev) for _ in scales] if automask else None
        args = (cfg, batch[("color", 0, 0)], [batch[("color", -1, 0)], batch[("color", 1, 0)]], batch[("K", 0)], batch[("inv_K", 0)], Ts,
                [batch[("ts", -1)], batch[("ts", 1)]], disps, flows, masks, noises)
        with torch.no_grad():
            t_f = ev_time(lambda: Fn.view_synthesis_sums(*args), reps)
        sums = Fn.view_synthesis_sums(*args)
        gs = torch.ones_like(sums) / (B * P)
        t_b = ev_time(lambda: sums.backward(gs, retain_graph=True), reps)
        fm = mode == 2
        fwd = B * sum(36 * P + 4 * P / 4**s + (16 * P / 4**s if fm else 0) + (8 * P if automask else 0) for s in scales)
        bwd = B * sum(36 * P + 8 * P / 4**s + (32 * P / 4**s if fm else 0) + (8 * P if automask else 0) for s in scales)
        print(f"warp {name:28s} fwd {t_f*1e3:8.1f} us ({fwd/t_f/1e6:7.1f} GB/s)   bwd {t_b*1e3:8.1f} us ({bwd/t_b/1e6:7.1f} GB/s)")


CONVS = [  # name, C0, C1, Cout, H, W (output), ksize, pad, act, up
    ("lite.1 240->112 @24x80 bil", 112, 128, 112, 24, 80, 3, "reflect", "elu", "bilinear"),
    ("lite.3 128->64 @48x160 bil", 64, 64, 64, 48, 160, 3, "reflect", "elu", "bilinear"),
    ("lite.5 32->32 @96x320 bil", 32, 0, 32, 96, 320, 3, "reflect", "elu", "bilinear"),
    ("motion L4 conv0 67->64 @96x320", 3, 64, 64, 96, 320, 3, "zero", "none", "none"),
    ("motion L4 conv1 64->64 @96x320", 64, 0, 64, 96, 320, 3, "zero", "none", "none"),
    ("motion L5 conv0 12->9 @192x640", 3, 9, 9, 192, 640, 3, "zero", "none", "none"),
    ("motion L4 redu 128->3 1x1", 64, 64, 3, 96, 320, 1, "zero", "none", "none"),
    ("motion L0 conv1 512->512 @6x20", 512, 0, 512, 6, 20, 3, "zero", "none", "none"),
    ("dispconv 32->1 @96x320", 32, 0, 1, 96, 320, 3, "reflect", "none", "none"),
]


def bench_conv(B, reps, only=None):
    dev = torch.device("cuda")
    for name, C0, C1, Cout, h, w, ks, pad, act, up in CONVS:
        if only and only not in name:
            continue
        h0, w0 = (h, w) if up == "none" else (h // 2, w // 2)
        x0 = torch.randn(B, C0, h0, w0, device=dev, requires_grad=True)
        x1 = torch.randn(B, C1, h, w, device=dev, requires_grad=True) if C1 else None
        wt = (torch.randn(Cout, C0 + C1, ks, ks, device=dev) * 0.05).requires_grad_(True)
        b = torch.zeros(Cout, device=dev, requires_grad=True)
        f = lambda: Fn.conv2d_fused(x0, wt, b, x1=x1, ksize=ks, pad=pad, act=act, up=up)
        with torch.no_grad():
            t_f = ev_time(f, reps)
        out = f()
        go = torch.randn_like(out)
        t_b = ev_time(lambda: out.backward(go, retain_graph=True), reps)
        flops = 2.0 * B * Cout * (C0 + C1) * ks * ks * h * w
        print(f"conv {name:34s} fwd {t_f*1e3:8.1f} us ({flops/t_f/1e9:6.1f} TF/s)  bwd(dgrad+wgrad) {t_b*1e3:8.1f} us ({2*flops/t_b/1e9:6.1f} TF/s)")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="all")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--only", default=None, help="substring of the convolution layer name")
    a = ap.parse_args()
    if a.what in ("warp", "all"):
        bench_warp(a.batch, a.reps)
    if a.what in ("warp0", "warp1", "warp2"):
        bench_warp(a.batch, a.reps, only=int(a.what[-1]))
    if a.what in ("conv", "all"):
        bench_conv(a.batch, a.reps, a.only)
