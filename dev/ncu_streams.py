"""One call of each round-2 encoder stream kernel at its stage-1 shape (bs32), for ncu --set full captures."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dynamo-depth_b200"))
from dd_b200 import functional as Fn

torch.manual_seed(0)
B = 32
for rep in range(2):
    qkv = torch.randn(B, 48 * 160, 192, device="cuda", requires_grad=True)
    temp = torch.ones(8, 1, 1, device="cuda", requires_grad=True)
    Fn.xca_core(qkv, temp, 8).sum().backward()
    x = torch.randn(B, 64, 48, 160, device="cuda", requires_grad=True)
    bn = torch.nn.BatchNorm2d(64).cuda().train()
    Fn.batch_norm_gelu(x, bn).sum().backward()
    xs = torch.randn(B, 64, 96, 320, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    Fn.bn_act_nhwc(xs, bn, "gelu").sum().backward()
    w = torch.randn(64, 1, 3, 3, device="cuda", requires_grad=True)
    Fn.dwconv3x3(x, w, 2).sum().backward()
    t = torch.randn(B, 48, 160, 64, device="cuda", requires_grad=True)
    g = torch.ones(64, device="cuda", requires_grad=True)
    Fn.layer_norm(t, g, g).sum().backward()
    torch.cuda.synchronize()
