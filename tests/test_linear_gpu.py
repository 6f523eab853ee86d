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


@pytest.mark.parametrize("B,C,HW", [(2, 64, 7680), (3, 224, 480), (2, 40, 100), (1, 7, 33)])
def test_layout_glue_kernels(B, C, HW):
    """dd_nchw_to_nhwc / dd_block_tail_* (csrc/lite_glue.cu) against the reference's permute / scale / add formulation."""
    from dd_b200.functional import block_tail, nchw_to_nhwc
    torch.manual_seed(B * 100 + C)
    H, W = (HW // 20, 20) if HW % 20 == 0 else (1, HW)
    x = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    y = torch.randn(B, H, W, C, device="cuda", requires_grad=True)
    gamma = torch.randn(C, device="cuda", requires_grad=True)
    scale = torch.tensor([0.0, 1.25, 1.25][:B], device="cuda")
    go = torch.randn(B, C, H, W, device="cuda")
    # transpose and its gradient
    t = nchw_to_nhwc(x)
    assert t.is_contiguous() and torch.equal(t, x.detach().permute(0, 2, 3, 1))
    gt = torch.randn_like(t)
    t.backward(gt)
    assert torch.equal(x.grad, gt.permute(0, 3, 1, 2))
    x.grad = None
    for use_gamma, use_scale in ((True, True), (True, False), (False, False)):
        x.grad = y.grad = gamma.grad = None
        out = block_tail(x, y, gamma if use_gamma else None, scale if use_scale else None)
        out.backward(go)
        got = (out.detach(), x.grad.clone(), y.grad.clone(), gamma.grad.clone() if use_gamma else None)
        x.grad = y.grad = gamma.grad = None
        z = (gamma * y if use_gamma else y).permute(0, 3, 1, 2)
        ref = x + (z * scale.view(-1, 1, 1, 1) if use_scale else z)
        ref.backward(go)
        assert torch.allclose(got[0], ref.detach(), rtol=1e-6, atol=1e-6)
        assert torch.equal(got[1], x.grad)
        assert torch.allclose(got[2], y.grad, rtol=1e-6, atol=1e-6)
        if use_gamma:
            assert _rel(got[3], gamma.grad.double()) < 1e-5


@pytest.mark.parametrize("train,drop", [(False, 0.0), (True, 0.0), (True, 0.2)])
def test_litemono_encoder_tc3x_matches_torch_fp32(train, drop, monkeypatch):
    """Same weights, same input, same RNG stream: encoder features and parameter gradients through the tcgen05 linear
    kernel + layout-glue kernels vs the reference formulation in plain torch fp32."""
    from networks import depth_encoder as de
    # fp32 cuDNN convolutions: with TF32 a one-ulp difference of an activation (the fused BatchNorm / LayerNorm kernels vs
    # ATen's) can round to a different TF32 operand and shows up as ~1e-4 in the next convolution's output
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    torch.manual_seed(11)
    enc = de.LiteMono(pretrained=False, drop_path_rate=drop).cuda()
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
