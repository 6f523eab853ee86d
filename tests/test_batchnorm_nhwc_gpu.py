"""dd_bn_act_nhwc_fwd / bwd (csrc/batchnorm_nhwc.cu) against nn.BatchNorm2d (+ residual) (+ ReLU / GELU) in float64 on
channels_last tensors -- the BasicBlock pattern of the ResNet trunks and the Lite-Mono stem's BNGELU.  Bound 1e-4
(north_star); held to 2e-5 of the tensor's scale.  Plus the fused BasicBlock path of ResnetEncoder against torchvision's."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

TOL = 2e-5


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


# (B, C, H, W, act, residual)
CASES = [(4, 64, 24, 40, "relu", False), (2, 64, 12, 20, "relu", True), (2, 128, 6, 10, "none", False), (3, 256, 3, 5, "relu", True),
         (2, 512, 2, 3, "relu", True), (2, 64, 17, 9, "gelu", False), (1, 8, 5, 7, "relu", True), (8, 16, 1, 1, "none", True)]


@pytest.mark.parametrize("B,C,H,W,act,use_res", CASES)
def test_bn_act_nhwc_matches_float64(B, C, H, W, act, use_res):
    from dd_b200.functional import bn_act_nhwc
    g = torch.Generator(device="cuda").manual_seed(B * 100 + C + H)
    cl = torch.channels_last
    x = (torch.randn(B, C, H, W, device="cuda", generator=g) * 1.5 + 2.0 * torch.randn(1, C, 1, 1, device="cuda", generator=g)).contiguous(memory_format=cl)
    res = torch.randn(B, C, H, W, device="cuda", generator=g).contiguous(memory_format=cl) if use_res else None
    gy = torch.randn(B, C, H, W, device="cuda", generator=g).contiguous(memory_format=cl)
    bn = nn.BatchNorm2d(C).cuda().train()
    with torch.no_grad():
        bn.weight.copy_(torch.randn(C, device="cuda", generator=g) * 0.5 + 1.0)
        bn.bias.copy_(torch.randn(C, device="cuda", generator=g) * 0.3)
    ref = nn.BatchNorm2d(C).cuda().double().train()
    ref.load_state_dict({k: (v.double() if v.is_floating_point() else v.clone()) for k, v in bn.state_dict().items()})

    xr = x.detach().double().requires_grad_(True)
    rr = res.detach().double().requires_grad_(True) if use_res else None
    yr = ref(xr)
    if use_res:
        yr = yr + rr
    yr = {"relu": torch.relu, "gelu": nn.functional.gelu, "none": lambda t: t}[act](yr)
    yr.backward(gy.double())

    xg = x.detach().clone().requires_grad_(True)
    rg = res.detach().clone().requires_grad_(True) if use_res else None
    y = bn_act_nhwc(xg, bn, act, residual=rg)
    assert y.is_contiguous(memory_format=cl) or y.numel() == y.shape[0] * y.shape[1]
    y.backward(gy)

    assert _rel(y.detach(), yr.detach()) < TOL
    assert _rel(xg.grad, xr.grad) < TOL
    if use_res:
        assert _rel(rg.grad, rr.grad) < TOL
    assert _rel(bn.weight.grad, ref.weight.grad) < TOL
    assert _rel(bn.bias.grad, ref.bias.grad) < TOL
    assert _rel(bn.running_mean, ref.running_mean) < TOL
    assert _rel(bn.running_var, ref.running_var) < TOL
    assert int(bn.num_batches_tracked) == 1


def test_resnet_encoder_fused_blocks_match_torchvision(monkeypatch):
    """ResnetEncoder (fused BN/ReLU/add, NHWC max-pool, NCHW hand-over) vs the same weights through torchvision's modules."""
    from networks.resnet_encoder import ResnetEncoder
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    torch.manual_seed(0)
    enc = ResnetEncoder(18, False, num_input_images=2).cuda().train()
    x = torch.rand(2, 6, 64, 96, device="cuda")
    res = {}
    for fused in (False, True):
        monkeypatch.setattr(ResnetEncoder, "channels_last", fused)
        monkeypatch.setattr(ResnetEncoder, "fused_blocks", fused)
        enc.zero_grad(set_to_none=True)
        feats = enc(x)
        sum(f.square().mean() for f in feats).backward()
        res[fused] = ([f.detach().clone() for f in feats], {n: p.grad.detach().clone() for n, p in enc.named_parameters() if p.grad is not None})
    for a, b in zip(res[True][0], res[False][0]):
        assert a.is_contiguous() and _rel(a, b) < 1e-4
    worst = max(_rel(res[True][1][n], gr) for n, gr in res[False][1].items())
    assert worst < 1e-4, worst
