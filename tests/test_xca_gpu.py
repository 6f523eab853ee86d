"""dd_xca_fwd / dd_xca_bwd (csrc/xca.cu) against the reference formulation of XCA.forward (networks/depth_encoder.py:63-83
between the qkv and proj layers) evaluated in float64: output, qkv gradient, temperature gradient.  Bound 1e-4 (north_star);
held to 2e-5 of the tensor's scale."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 2e-5


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def ref_core(qkv, temp, heads):
    B, N, C3 = qkv.shape
    C = C3 // 3
    t = qkv.reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 4, 1)
    q, k, v = F.normalize(t[0], dim=-1), F.normalize(t[1], dim=-1), t[2]
    attn = ((q @ k.transpose(-2, -1)) * temp).softmax(dim=-1)
    return (attn @ v).permute(0, 3, 1, 2).reshape(B, N, C)


# (B, N, C, heads): the three encoder widths (d = 8, 16, 28), token counts that are not multiples of the 16-token tile
CASES = [(2, 480, 64, 8), (3, 77, 64, 8), (2, 240, 128, 8), (2, 61, 224, 8), (1, 5, 32, 4), (2, 1000, 64, 8)]


@pytest.mark.parametrize("B,N,C,heads", CASES)
def test_xca_core_matches_float64(B, N, C, heads):
    from dd_b200.functional import xca_core
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + N + C)
    qkv = torch.randn(B, N, 3 * C, device="cuda", generator=g, requires_grad=True)
    temp = (torch.rand(heads, 1, 1, device="cuda", generator=g) * 4 + 0.5).requires_grad_(True)   # sharpened attention rows
    gy = torch.randn(B, N, C, device="cuda", generator=g)
    y = xca_core(qkv, temp, heads)
    y.backward(gy)
    qr, tr = qkv.detach().double().requires_grad_(True), temp.detach().double().requires_grad_(True)
    yr = ref_core(qr, tr, heads)
    yr.backward(gy.double())
    assert _rel(y.detach(), yr.detach()) < TOL
    assert _rel(qkv.grad, qr.grad) < TOL
    assert _rel(temp.grad, tr.grad) < TOL


def test_xca_module_uses_kernel_and_matches_torch_mode():
    """The XCA module through the fused core vs its reference formulation (EncoderLinear.mode = "torch")."""
    from networks import depth_encoder as de
    torch.manual_seed(3)
    m = de.XCA(64, num_heads=8, qkv_bias=True).cuda()
    with torch.no_grad():
        m.temperature.uniform_(0.5, 3.0)
    x = torch.randn(2, 240, 64, device="cuda")
    res = {}
    for mode in ("torch", "tc3x"):
        de.EncoderLinear.mode = mode
        m.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        y = m(xi)
        y.square().sum().backward()
        res[mode] = (y.detach(), xi.grad, m.temperature.grad.clone(), m.qkv.weight.grad.clone())
    de.EncoderLinear.mode = "tc3x"
    for a, b in zip(res["tc3x"], res["torch"]):
        assert _rel(a, b) < 1e-4


def test_xca_rejects_bad_inputs():
    from dd_b200 import _lib as L
    from dd_b200.functional import xca_core
    with pytest.raises(L.DynamoB200Error):
        xca_core(torch.randn(1, 8, 96), torch.ones(4, 1, 1), 4)                                    # CPU tensors
    with pytest.raises(L.DynamoB200Error):
        xca_core(torch.randn(1, 8, 3 * 48, device="cuda"), torch.ones(4, 1, 1, device="cuda"), 4)   # d = 12 not built
