"""dd_maxpool3x3s2_nhwc_fwd / bwd (csrc/pool.cu) against F.max_pool2d(3, 2, 1): values bit-exact, gradients bit-exact
(including ties after ReLU, where the first maximum in scan order takes the gradient as in ATen), odd sizes."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,C,H,W,relu", [(2, 64, 24, 40, True), (1, 8, 7, 9, False), (3, 4, 1, 5, True), (2, 12, 6, 6, True), (1, 4, 2, 2, False)])
def test_maxpool_matches_torch(B, C, H, W, relu):
    from dd_b200.functional import maxpool3x3s2
    g = torch.Generator(device="cuda").manual_seed(B + C + H + W)
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    if relu:
        x = x.relu()   # many exact ties at 0
    x = x.contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = maxpool3x3s2(xa), F.max_pool2d(xb, 3, 2, 1)
    assert ya.shape == yb.shape and ya.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(ya, yb)
    gy = torch.randn(yb.shape, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    ya.backward(gy)
    yb.backward(gy)
    assert torch.allclose(xa.grad, xb.grad, rtol=0, atol=1e-6)   # sums of up to four gradients: order of addition only


def test_maxpool_nan_and_refusals():
    from dd_b200 import _lib as L
    from dd_b200.functional import maxpool3x3s2
    x = torch.randn(1, 4, 6, 6, device="cuda").contiguous(memory_format=torch.channels_last)
    x[0, 1, 2, 3] = float("nan")
    a, b = maxpool3x3s2(x), F.max_pool2d(x, 3, 2, 1)
    assert torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(a.nan_to_num(7.0), b.nan_to_num(7.0))
    with pytest.raises(L.DynamoB200Error):
        maxpool3x3s2(torch.randn(1, 4, 6, 6))
    with pytest.raises(L.DynamoB200Error):
        maxpool3x3s2(torch.randn(1, 6, 6, 6, device="cuda"))   # C % 4 != 0


def test_to_nchw_handover_roundtrip():
    """channels_last -> NCHW hand-over of the ResNet features (dd_nhwc_to_nchw) and its gradient back in channels_last."""
    from dd_b200.functional import to_nchw
    torch.manual_seed(1)
    for shape in [(2, 64, 12, 20), (3, 20, 5, 7), (1, 512, 3, 10)]:
        x = torch.randn(*shape, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = to_nchw(x)
        assert y.is_contiguous() and torch.equal(y, x.detach().contiguous())
        g = torch.randn(*shape, device="cuda")
        y.backward(g)
        assert torch.equal(x.grad, g) and x.grad.is_contiguous(memory_format=torch.channels_last)
    z = torch.randn(2, 8, 4, 4, device="cuda")
    assert to_nchw(z) is z   # already NCHW-contiguous: passes through


def test_ground_plane_index_staging_ring():
    """GroundPlane._stage_indices: pinned ring + non-blocking copies deliver every draw intact even when the ring wraps while
    earlier copies are still queued behind GPU work."""
    import numpy as np
    from tools import GroundPlane
    gp = GroundPlane()
    rng = np.random.default_rng(0)
    busy = torch.randn(4096, 4096, device="cuda")
    draws, staged = [], []
    for i in range(3 * GroundPlane._PIN_SLOTS + 1):
        busy = busy @ busy.clamp(-1e-3, 1e-3)           # keep the stream occupied so that copies queue up
        a = rng.integers(0, 1000, size=(4, 125))
        draws.append(a)
        staged.append(gp._stage_indices(a, torch.device("cuda")))
    torch.cuda.synchronize()
    for a, t in zip(draws, staged):
        assert np.array_equal(t.cpu().numpy(), a)
