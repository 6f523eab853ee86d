"""Stand-alone timing of the fused BatchNorm(+GELU) kernels against torch / cuDNN on the stem shape (bs32, 64 x 96 x 320)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dynamo-depth_b200"))
from dd_b200.functional import batch_norm_gelu

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for (B, C, H, W, gelu) in [(32, 64, 96, 320, True), (32, 64, 48, 160, False), (32, 128, 24, 80, False), (32, 224, 12, 40, False)]:
    x = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    gy = torch.randn_like(x)
    bn = torch.nn.BatchNorm2d(C).cuda().train()
    mb = x.numel() * 4 / 1e6
    def ours_f(): return batch_norm_gelu(x, bn, gelu=gelu)
    def torch_f():
        y = bn(x)
        return torch.nn.functional.gelu(y) if gelu else y
    res = {}
    for name, f in (("ours", ours_f), ("torch", torch_f)):
        tf = timed(f)
        y = f()
        def b(): torch.autograd.grad(y, [x, bn.weight, bn.bias], gy, retain_graph=True)
        tb = timed(b)
        res[name] = (tf, tb)
    print(f"bn{'+gelu' if gelu else ''} {B}x{C}x{H}x{W} ({mb:.0f} MB): fwd ours {res['ours'][0]:.0f} us ({3*mb/res['ours'][0]*1e-3*1e3:.0f} GB/s) torch {res['torch'][0]:.0f} us | "
          f"bwd ours {res['ours'][1]:.0f} us ({5*mb/res['ours'][1]*1e-3*1e3:.0f} GB/s) torch {res['torch'][1]:.0f} us")
