"""GPU parity of the smoothness / motion-sparsity kernels against the CPU oracle."""
import pytest
import torch

from oracle import view_synthesis as vs

pytestmark = pytest.mark.gpu


def test_smooth_sums_and_grads():
    from dd_b200 import functional as Fn

    g = torch.Generator().manual_seed(3)
    shapes = [(2, 1, 16, 24), (2, 3, 16, 24), (2, 1, 8, 12), (3, 3, 32, 64)]
    norms = [True, False, False, False]
    inps = [torch.rand(s, generator=g) + 0.05 for s in shapes]
    imgs = [torch.rand(s[0], 3, s[2], s[3], generator=g) for s in shapes]
    imgs[2] = None
    # oracle
    o_inps = [t.clone().requires_grad_(True) for t in inps]
    o_vals = []
    for t, im, n in zip(o_inps, imgs, norms):
        x = t / (t.mean(2, True).mean(3, True) + 1e-7) if n else t
        o_vals.append(vs.smooth_loss(x, im))
    wts = torch.tensor([0.3, 1.7, -0.9, 2.2])
    (torch.stack(o_vals) * wts).sum().backward()
    # cuda
    c_inps = [t.cuda().requires_grad_(True) for t in inps]
    sums = Fn.smooth_sums(c_inps, [im.cuda() if im is not None else None for im in imgs], norms)
    vals = Fn.smooth_means(sums, shapes)
    (vals * wts.cuda()).sum().backward()
    for a, e in zip(vals.tolist(), o_vals):
        assert a == pytest.approx(float(e), rel=1e-4)
    for a, e in zip(c_inps, o_inps):
        scale = e.grad.abs().max().item()
        assert (a.grad.cpu() - e.grad).abs().max().item() <= 1e-4 * scale + 1e-9


@pytest.mark.parametrize("empty_image", [False, True])
def test_motion_sparsity(empty_image):
    from dd_b200 import functional as Fn

    g = torch.Generator().manual_seed(5)
    B, h, w = 3, 12, 20
    mag = torch.rand(B, h, w, generator=g)
    if empty_image:
        mag[1] = 10.0   # no static pixel in image 1 -> the term is skipped (Trainer.py:398)
    prob = (4 * torch.randn(B, 1, h, w, generator=g)).requires_grad_(True)
    static = (mag < mag.mean()).unsqueeze(1)
    if bool(torch.all(static.sum((1, 2, 3)) > 0)):
        ref = vs.softplus_mean(prob[static])
        ref.backward()
        ref_v, ref_g = float(ref), prob.grad.clone()
    else:
        ref_v, ref_g = 0.0, torch.zeros_like(prob)
    pc = prob.detach().cuda().requires_grad_(True)
    magc = mag.cuda()
    out = Fn.motion_sparsity(magc, magc.sum().reshape(1), pc)
    out.backward()
    assert float(out) == pytest.approx(ref_v, rel=1e-4, abs=1e-7)
    assert (pc.grad.cpu() - ref_g).abs().max().item() <= 1e-4 * (ref_g.abs().max().item() + 1e-9) + 1e-9
