"""Cost of the XCA attention core (everything between the qkv and proj linear layers) at the three Lite-Mono stage shapes."""
import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dynamo-depth_b200"))

def core(qkv, temp, heads):
    B, N, C3 = qkv.shape
    C = C3 // 3
    t = qkv.reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 4, 1)
    q, k, v = F.normalize(t[0], dim=-1), F.normalize(t[1], dim=-1), t[2]
    attn = ((q @ k.transpose(-2, -1)) * temp).softmax(dim=-1)
    return (attn @ v).permute(0, 3, 1, 2).reshape(B, N, C)

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

tot_f = tot_b = 0
for (N, C) in [(48 * 160, 64), (24 * 80, 128), (12 * 40, 224)]:
    qkv = torch.randn(32, N, 3 * C, device="cuda", requires_grad=True)
    temp = torch.ones(8, 1, 1, device="cuda", requires_grad=True)
    g = torch.randn(32, N, C, device="cuda")
    tf = timed(lambda: core(qkv, temp, 8))
    y = core(qkv, temp, 8)
    tb = timed(lambda: torch.autograd.grad(y, [qkv, temp], g, retain_graph=True))
    print(f"N={N} C={C}: fwd {tf:.0f} us bwd {tb:.0f} us  (qkv {qkv.numel()*4/1e6:.0f} MB)")
    tot_f += tf; tot_b += tb
print(f"per step: 3 x fwd {3*tot_f/1e3:.2f} ms + 1 x bwd {tot_b/1e3:.2f} ms")
from dd_b200.functional import xca_core
tot_f = tot_b = 0
for (N, C) in [(48 * 160, 64), (24 * 80, 128), (12 * 40, 224)]:
    qkv = torch.randn(32, N, 3 * C, device="cuda", requires_grad=True)
    temp = torch.ones(8, 1, 1, device="cuda", requires_grad=True)
    g = torch.randn(32, N, C, device="cuda")
    tf = timed(lambda: xca_core(qkv, temp, 8))
    y = xca_core(qkv, temp, 8)
    tb = timed(lambda: torch.autograd.grad(y, [qkv, temp], g, retain_graph=True))
    print(f"fused N={N} C={C}: fwd {tf:.0f} us bwd {tb:.0f} us")
    tot_f += tf; tot_b += tb
print(f"fused per step: 3 x fwd {3*tot_f/1e3:.2f} ms + 1 x bwd {tot_b/1e3:.2f} ms")
