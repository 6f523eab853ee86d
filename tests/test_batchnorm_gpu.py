"""dd_bn_gelu_fwd / dd_bn_gelu_bwd (csrc/batchnorm.cu) against nn.BatchNorm2d (+ nn.GELU) evaluated in float64: output,
input / affine gradients, saved and running statistics, num_batches_tracked -- the training-mode semantics of the reference's
BNGELU (networks/depth_encoder.py:113-122) and DilatedConv.bn1 (:194,:208).  Tolerance: north_star's 1e-4 is the bound; the
kernels are held to 2e-5 of the tensor's scale (measured <= 3e-6)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

TOL = 2e-5


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


# (B, C, H, W, gelu): stem-like, vector tail (HW % 4 != 0), single pixel rows, more channels than chunks, large offset mean
CASES = [(4, 64, 24, 40, True), (3, 5, 7, 9, True), (2, 224, 12, 40, False), (8, 16, 33, 17, False), (2, 3, 64, 96, True),
         (1, 7, 1, 3, False)]


@pytest.mark.parametrize("B,C,H,W,gelu", CASES)
def test_bn_gelu_matches_float64(B, C, H, W, gelu):
    from dd_b200.functional import batch_norm_gelu
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + C * 10 + H)
    x = (torch.randn(B, C, H, W, device="cuda", generator=g) * 1.7 + 3.0 * torch.randn(1, C, 1, 1, device="cuda", generator=g))
    gy = torch.randn(B, C, H, W, device="cuda", generator=g)
    bn = nn.BatchNorm2d(C, eps=1e-5).cuda().train()
    with torch.no_grad():
        bn.weight.copy_(torch.randn(C, device="cuda", generator=g) * 0.5 + 1.0)
        bn.bias.copy_(torch.randn(C, device="cuda", generator=g) * 0.3)
        bn.running_mean.copy_(torch.randn(C, device="cuda", generator=g))
        bn.running_var.copy_(torch.rand(C, device="cuda", generator=g) + 0.5)
    ref = nn.BatchNorm2d(C, eps=1e-5).cuda().double().train()
    ref.load_state_dict({k: (v.double() if v.is_floating_point() else v.clone()) for k, v in bn.state_dict().items()})

    xr = x.detach().double().requires_grad_(True)
    yr = ref(xr)
    if gelu:
        yr = nn.functional.gelu(yr)
    yr.backward(gy.double())

    xg = x.detach().clone().requires_grad_(True)
    y = batch_norm_gelu(xg, bn, gelu=gelu)
    y.backward(gy)

    assert _rel(y.detach(), yr.detach()) < TOL
    assert _rel(xg.grad, xr.grad) < TOL
    assert _rel(bn.weight.grad, ref.weight.grad) < TOL
    assert _rel(bn.bias.grad, ref.bias.grad) < TOL
    assert _rel(bn.running_mean, ref.running_mean) < TOL
    assert _rel(bn.running_var, ref.running_var) < TOL
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked) == 1


def test_bn_eval_mode_and_frozen_affine():
    """Eval mode keeps torch's running-statistics path; frozen affine parameters still give the input gradient."""
    from dd_b200.functional import batch_norm_gelu
    torch.manual_seed(5)
    bn = nn.BatchNorm2d(8).cuda()
    x = torch.randn(2, 8, 6, 10, device="cuda")
    bn.eval()
    assert torch.equal(batch_norm_gelu(x, bn, gelu=True), nn.functional.gelu(bn(x)))
    bn.train()
    for p in bn.parameters():
        p.requires_grad_(False)
    xg = x.clone().requires_grad_(True)
    gy = torch.randn_like(x)
    batch_norm_gelu(xg, bn).backward(gy)
    xr = x.double().requires_grad_(True)
    ref = nn.BatchNorm2d(8).cuda().double().train()
    ref(xr).backward(gy.double())
    assert _rel(xg.grad, xr.grad) < TOL


def test_bn_rejects_cpu():
    from dd_b200 import _lib as L
    from dd_b200.functional import batch_norm_gelu
    bn = nn.BatchNorm2d(4).train()
    with pytest.raises(L.DynamoB200Error):
        batch_norm_gelu(torch.randn(2, 4, 8, 8), bn)
