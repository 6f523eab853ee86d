"""dd_dwconv3x3_fwd / dd_dwconv3x3_wgrad (csrc/dwconv.cu) against F.conv2d(groups=C) in float64: the depth-wise dilated
convolution of the reference's CDilated (networks/depth_encoder.py:148-168).  Bound 1e-4 (north_star); held to 1e-5 of the
tensor's scale (9-term sums in fp32)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


# (B, C, H, W, dilation): every built dilation, maps smaller than the dilated footprint, one-vector rows
CASES = [(2, 8, 12, 40, 1), (2, 8, 12, 40, 2), (3, 5, 9, 16, 3), (2, 4, 7, 12, 4), (2, 6, 12, 40, 6), (1, 3, 3, 4, 6), (1, 2, 1, 4, 1),
         (2, 64, 48, 160, 2)]


@pytest.mark.parametrize("B,C,H,W,d", CASES)
def test_dwconv_matches_float64(B, C, H, W, d):
    from dd_b200.functional import dwconv3x3
    g = torch.Generator(device="cuda").manual_seed(B + 10 * C + 100 * d)
    x = torch.randn(B, C, H, W, device="cuda", generator=g, requires_grad=True)
    w = torch.randn(C, 1, 3, 3, device="cuda", generator=g, requires_grad=True)
    gy = torch.randn(B, C, H, W, device="cuda", generator=g)
    y = dwconv3x3(x, w, d)
    y.backward(gy)
    xr, wr = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, 1, d, d, C)
    yr.backward(gy.double())
    assert _rel(y.detach(), yr.detach()) < TOL
    assert _rel(x.grad, xr.grad) < TOL
    assert _rel(w.grad, wr.grad) < TOL


def test_dwconv_rejects_bad_inputs():
    from dd_b200 import _lib as L
    from dd_b200.functional import dwconv3x3
    with pytest.raises(L.DynamoB200Error):
        dwconv3x3(torch.randn(1, 2, 4, 8), torch.randn(2, 1, 3, 3), 1)                                   # CPU tensors
    with pytest.raises(L.DynamoB200Error):
        dwconv3x3(torch.randn(1, 2, 4, 6, device="cuda"), torch.randn(2, 1, 3, 3, device="cuda"), 1)     # W % 4 != 0
    with pytest.raises(L.DynamoB200Error):
        dwconv3x3(torch.randn(1, 2, 4, 8, device="cuda"), torch.randn(2, 1, 3, 3, device="cuda"), 5)     # dilation not built
