"""Bring-up check of dd_linear_fwd / dd_linear_bwd (csrc/linear_tc.cu) against float64 matmuls on the GPU.

    python dev/linear_tc_test.py [--bench]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dynamo-depth_b200"))
from dd_b200 import functional as Fn  # noqa: E402


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check(M, K, N, bias=True, what=("fwd", "dx", "dw")):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + K * 3 + N)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    gy = torch.randn(M, N, device="cuda", generator=g)
    x.requires_grad_(True), w.requires_grad_(True)
    if bias:
        b.requires_grad_(True)
    y = Fn.linear(x, w, b)
    torch.cuda.synchronize()
    yr = x.detach().double() @ w.detach().double().t() + (b.detach().double() if bias else 0)
    out = {"fwd": rel(y.detach(), yr)}
    y32 = torch.nn.functional.linear(x.detach(), w.detach(), b.detach() if bias else None)
    out["fwd_torch32"] = rel(y32, yr)
    y.backward(gy)
    torch.cuda.synchronize()
    out["dx"] = rel(x.grad, gy.double() @ w.detach().double())
    out["dw"] = rel(w.grad, gy.double().t() @ x.detach().double())
    if bias:
        out["db"] = rel(b.grad, gy.double().sum(0))
    print(f"M={M:7d} K={K:5d} N={N:5d} " + " ".join(f"{k}={v:.2e}" for k, v in out.items()), flush=True)
    return max(v for k, v in out.items() if k != "fwd_torch32")


def bench(M, K, N, reps=10):
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda")
    b = torch.randn(N, device="cuda")
    gy = torch.randn(M, N, device="cuda")
    lib = Fn.L.load()
    y = torch.empty(M, N, device="cuda")
    gx = torch.empty(M, K, device="cuda")
    gw = torch.empty(N, K, device="cuda")
    gb = torch.empty(N, device="cuda")
    st = Fn._stream()

    def t(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    P = Fn.L.ptr
    ws = torch.empty(lib.dd_linear_workspace_bytes(M, K, N), dtype=torch.uint8, device="cuda")
    res = {
        "fwd": t(lambda: lib.dd_linear_fwd(P(x), P(w), P(b), M, K, N, P(y), P(ws), ws.numel(), st)),
        "dx": t(lambda: lib.dd_linear_bwd(P(x), P(w), P(gy), M, K, N, P(gx), None, None, P(ws), ws.numel(), st)),
        "dw": t(lambda: lib.dd_linear_bwd(P(x), P(w), P(gy), M, K, N, None, P(gw), P(gb), P(ws), ws.numel(), st)),
        "t32_fwd": t(lambda: torch.nn.functional.linear(x, w, b)),
        "t32_dx": t(lambda: gy @ w),
        "t32_dw": t(lambda: gy.t() @ x),
    }
    flop = 2.0 * M * K * N
    io = 4.0 * (M * K + M * N)
    print(f"M={M:7d} K={K:5d} N={N:5d} " + " ".join(f"{k}={v * 1e3:7.1f}us" for k, v in res.items()) +
          f" | fwd {flop / res['fwd'] / 1e9:6.1f} TF/s {io / res['fwd'] / 1e6:6.0f} GB/s", flush=True)


if __name__ == "__main__":
    torch.manual_seed(0)
    worst = 0.0
    for shape in [(128, 32, 32), (128, 64, 64), (256, 64, 384), (1000, 64, 192), (4096, 384, 64), (3000, 224, 1344),
                  (5000, 1344, 224), (2048, 128, 768), (2048, 768, 128), (777 * 4, 224, 672), (20000, 64, 384)]:
        worst = max(worst, check(*shape))
    print("worst relative error", worst)
    if "--bench" in sys.argv:
        for shape in [(245760, 64, 384), (245760, 384, 64), (61440, 128, 768), (61440, 768, 128), (15360, 224, 1344),
                      (15360, 1344, 224), (15360, 224, 672), (15360, 224, 224)]:
            bench(*shape)
    sys.exit(0 if worst < 2e-5 else 1)
