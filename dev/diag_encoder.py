"""Per-module output difference of the Lite-Mono encoder between EncoderLinear.mode = torch and tc3x (train mode)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dynamo-depth_b200"))
from networks import depth_encoder as de
torch.manual_seed(11)
enc = de.LiteMono(pretrained=False, drop_path_rate=0.0).cuda().train()
x = torch.rand(2, 3, 96, 160, device="cuda")
outs = {}
def hook(name):
    def f(m, i, o):
        if torch.is_tensor(o): outs.setdefault(name, []).append(o.detach().clone())
    return f
for n, m in enc.named_modules():
    if n: m.register_forward_hook(hook(n))
for mode in ("torch", "tc3x"):
    de.EncoderLinear.mode = mode
    torch.manual_seed(5)
    with torch.no_grad(): enc(x)
worst = 0
for n, (a, b) in ((k, v) for k, v in outs.items() if len(v) == 2):
    r = float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
    if r > 2e-6 and r > worst * 1.5:
        worst = r
        print(f"{n:50s} {tuple(a.shape)} rel {r:.3e}")
