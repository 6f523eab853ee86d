"""dd_linear_fwd / dd_linear_bwd (csrc/linear_tc.cu: tcgen05 3xTF32, fp32 accumulation in TMEM) against float64
matmuls, and the Lite-Mono encoder through that kernel against torch's fp32 matmul path.

Tolerance: 1e-4 relative fp32 is the north-star bound; the kernel itself is held to 2e-5 of the largest output
magnitude (observed <= 1.2e-5 at reduction length 1344: the tensor core's fp32 accumulator truncates, so the error
grows slowly with the reduction length; torch's SIMT fp32 GEMM measures ~1e-6 on the same problems)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-5


def _rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


# (M, K, N, bias): tile-aligned, ragged rows, N tail tile, K not a multiple of the 32-wide K block, single tile
SHAPES = [
    (128, 32, 32, True), (256, 64, 384, True), (1000, 64, 192, True), (4096, 384, 64, True), (3000, 224, 1344, True),
    (5000, 1344, 224, True), (2048, 128, 768, False), (3108, 224, 672, True), (516, 36, 100, True), (4, 8, 4, True),
    (130, 260, 516, False),
]


@pytest.mark.parametrize("M,K,N,bias", SHAPES)
def test_linear_matches_float64(M, K, N, bias):
    from dd_b200.functional import linear
    g = torch.Generator(device="cuda").manual_seed(M * 7 + K * 3 + N)
    x = torch.randn(M, K, device="cuda", generator=g, requires_grad=True)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).requires_grad_(True)
    b = torch.randn(N, device="cuda", generator=g, requires_grad=True) if bias else None
    gy = torch.randn(M, N, device="cuda", generator=g)
    y = linear(x, w, b)
    y.backward(gy)
    xd, wd, gd = x.detach().double(), w.detach().double(), gy.double()
    assert _rel(y.detach(), xd @ wd.t() + (b.detach().double() if bias else 0)) < TOL
    assert _rel(x.grad, gd @ wd) < TOL
    assert _rel(w.grad, gd.t() @ xd) < TOL
    if bias:
        assert _rel(b.grad, gd.sum(0)) < TOL


def test_linear_leading_dims_and_partial_grads():
    """(B, H, W, C) inputs as the encoder blocks pass them; frozen weight (input gradient only)."""
    from dd_b200.functional import linear
    torch.manual_seed(3)
    x = torch.randn(2, 12, 40, 64, device="cuda", requires_grad=True)
    w = torch.randn(384, 64, device="cuda") / 8
    b = torch.randn(384, device="cuda")
    y = linear(x, w, b)
    assert y.shape == (2, 12, 40, 384)
    y.square().sum().backward()
    ref = torch.nn.functional.linear(x.detach().double(), w.double(), b.double())
    assert _rel(y.detach(), ref) < TOL
    assert _rel(x.grad, (2 * ref) @ w.double()) < TOL


def test_linear_rejects_cpu_and_bad_shapes():
    from dd_b200._lib import DynamoB200Error
    from dd_b200.functional import linear
    with pytest.raises(DynamoB200Error):
        linear(torch.randn(8, 8), torch.randn(8, 8))
    with pytest.raises(DynamoB200Error):   # K % 4 != 0
        linear(torch.randn(8, 6, device="cuda"), torch.randn(8, 6, device="cuda"))


@pytest.mark.parametrize("train", [False, True])
def test_litemono_encoder_tc3x_matches_torch_fp32(train):
    """Same weights, same input: encoder features and parameter gradients through the tcgen05 kernel vs torch fp32."""
    from networks import depth_encoder as de
    torch.manual_seed(11)
    enc = de.LiteMono(pretrained=False, drop_path_rate=0.0).cuda()
    enc.train(train)
    x = torch.rand(2, 3, 96, 160, device="cuda")
    res = {}
    for mode in ("torch", "tc3x"):
        de.EncoderLinear.mode = mode
        torch.manual_seed(5)
        enc.zero_grad(set_to_none=True)
        feats = enc(x)
        sum(f.square().mean() for f in feats).backward()
        res[mode] = ([f.detach().clone() for f in feats],
                     {n: p.grad.detach().clone() for n, p in enc.named_parameters() if p.grad is not None})
    de.EncoderLinear.mode = "tc3x"
    for a, b in zip(res["tc3x"][0], res["torch"][0]):
        assert _rel(a, b.double()) < 1e-4
    worst = max(_rel(res["tc3x"][1][n], g.double()) for n, g in res["torch"][1].items())
    assert worst < 1e-4, worst
