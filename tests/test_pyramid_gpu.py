"""Input colour pyramid kernel (dd_pyramid_half_fwd) against the reference's own call: torchvision
Resize(BICUBIC, antialias=True) + clamp (Trainer.py:80, 729-734), evaluated on the CPU in fp32."""
import pytest
import torch
import torchvision


pytestmark = pytest.mark.gpu


def _reference_chain(x, levels):
    out = [x]
    for s in range(1, levels + 1):
        h, w = x.shape[-2] >> s, x.shape[-1] >> s
        rs = torchvision.transforms.Resize((h, w), interpolation=torchvision.transforms.InterpolationMode.BICUBIC, antialias=True)
        out.append(torch.clamp(rs(out[-1]), 0, 1))
    return out


@pytest.mark.parametrize("shape", [(2, 3, 192, 640), (3, 3, 64, 96), (1, 3, 32, 36), (1, 1, 6, 8), (2, 3, 2, 2)])
def test_pyramid_half_matches_torchvision(shape):
    from dd_b200 import functional as Fn

    g = torch.Generator().manual_seed(sum(shape))
    x = torch.rand(shape, generator=g)
    x[..., : shape[-2] // 2, :] *= 1.6          # saturate part of the image so that the clamp is exercised
    x = x.clamp(0, 1.3) - 0.1
    levels = 2 if min(shape[-2:]) >= 8 else 1
    ref = _reference_chain(x, levels)
    cur = x.cuda()
    for s in range(1, levels + 1):
        cur = Fn.pyramid_half(cur)
        assert cur.shape == ref[s].shape
        err = (cur.cpu() - ref[s]).abs().max().item()
        assert err <= 2e-6, (s, err)             # fp32 rounding of 64 products; weights are normalised in fp32 as in ATen


def test_pyramid_rejects_cpu_and_odd_sizes():
    from dd_b200 import _lib as L
    from dd_b200 import functional as Fn

    with pytest.raises(L.DynamoB200Error):
        Fn.pyramid_half(torch.rand(1, 3, 8, 8))
    with pytest.raises(L.DynamoB200Error):
        Fn.pyramid_half(torch.rand(1, 3, 7, 8, device="cuda"))


def test_trainer_pyramid_uses_kernel_and_matches_host_chain():
    import options
    from Trainer import Trainer
    from dd_b200 import _lib as L
    from dd_b200 import synthetic

    opt = options.DynamoOptions().parse(args=["-d", "waymo", "--depth_model", "litemono", "--weights_init", "scratch", "--height", "64",
                                              "--width", "96", "-b", "2"])
    opt.cuda_ids, opt.local_rank, opt.ddp = [0], 0, False
    tr = Trainer(opt)
    batch = synthetic.make_batch(opt, 5)
    host = dict(batch)
    for s in opt.scales:
        if s != 0:
            host[("color", 0, s)] = torch.clamp(tr.resize[s](host[("color", 0, s - 1)]), 0, 1)
    dev = {k: v.cuda() for k, v in batch.items()}
    n0 = L.load().dd_launch_count()
    tr.apply_img_resize(dev)
    assert L.load().dd_launch_count() - n0 == len([s for s in opt.scales if s != 0])
    for s in opt.scales:
        assert (dev[("color", 0, s)].cpu() - host[("color", 0, s)]).abs().max().item() <= 2e-6


def test_prefetched_batch_equals_direct_batch():
    """Trainer.prefetch (copy stream + event) must hand process_batch the same tensors as the in-line copy."""
    import options
    from Trainer import Trainer
    from dd_b200 import synthetic

    opt = options.DynamoOptions().parse(args=["-d", "waymo", "--depth_model", "litemono", "--weights_init", "scratch", "--height", "64",
                                              "--width", "96", "-b", "2"])
    opt.cuda_ids, opt.local_rank, opt.ddp = [0], 0, False
    tr = Trainer(opt)
    batch = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in synthetic.make_batch(opt, 9).items()}
    direct = dict(batch)
    tr.process_inputs(direct)
    staged = tr.prefetch(dict(batch))
    assert "__ready__" in staged
    tr.process_inputs(staged)
    assert "__ready__" not in staged
    torch.cuda.synchronize()
    assert set(staged) == set(direct)
    for k, v in direct.items():
        if torch.is_tensor(v):
            assert staged[k].device == v.device and torch.equal(staged[k], v), k
