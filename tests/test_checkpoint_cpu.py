"""Training-state checkpoint (dd_b200/checkpoint.py) on the CPU tier: a run that is saved, rebuilt from
scratch and resumed reproduces the uninterrupted run bit for bit (Adam moments, LR schedule, RNG streams, counters)."""
import numpy as np
import pytest
import torch

from dd_b200 import checkpoint as ckpt


def _make(seed=0):
    torch.manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Linear(6, 12), torch.nn.Dropout(0.3), torch.nn.Linear(12, 3))
    opt = torch.optim.Adam(net.parameters(), 1e-2)
    sched = torch.optim.lr_scheduler.StepLR(opt, 2, 0.5)
    return net, opt, sched


def _epoch(net, opt, sched, steps=3):
    net.train()
    for _ in range(steps):
        x = torch.randn(5, 6) + float(np.random.rand())      # consumes the torch and the numpy streams
        net(x).square().mean().backward()
        opt.step()
        opt.zero_grad()
    sched.step()


def test_resume_reproduces_uninterrupted_run(tmp_path):
    net, opt, sched = _make()
    np.random.seed(3)
    for _ in range(4):
        _epoch(net, opt, sched)
    want = [p.detach().clone() for p in net.parameters()]

    net, opt, sched = _make()
    np.random.seed(3)
    for epoch in range(2):
        _epoch(net, opt, sched)
    ckpt.save_state(str(tmp_path), ckpt.pack_state("mask_init", 1, 6, 11, opt, sched, [1, 1, 5, 20]))
    torch.save(net.state_dict(), tmp_path / "net.pth")

    net2, opt2, sched2 = _make(seed=99)                        # a fresh process: different init, different RNG position
    np.random.seed(1234)
    torch.randn(7)
    state = ckpt.load_state(str(tmp_path))
    net2.load_state_dict(torch.load(tmp_path / "net.pth"))
    ckpt.apply_state(state, opt2, sched2)
    assert (state["phase_name"], state["epoch"], state["step"], state["g_step"]) == ("mask_init", 1, 6, 11)
    assert ckpt.resume_point(state, [1, 1, 5, 20]) == (2, 2)
    assert sched2.get_last_lr() == sched.get_last_lr()
    for _ in range(2):
        _epoch(net2, opt2, sched2)
    for a, b in zip(net2.parameters(), want):
        assert torch.equal(a.detach(), b)


def test_resume_point_rolls_over_to_the_next_phase():
    _, opt, sched = _make()
    st = ckpt.pack_state("disp_init", 0, 10, 10, opt, sched, [1, 0, 5, 20], with_rng=False)
    assert ckpt.resume_point(st, [1, 0, 5, 20]) == (1, 0)       # disp_init finished: Trainer.train skips empty phases itself
    st = ckpt.pack_state("fine_tune", 19, 10, 10, opt, sched, [1, 1, 5, 20], with_rng=False)
    assert ckpt.resume_point(st, [1, 1, 5, 20]) == (4, 0)       # nothing left to run
    with pytest.raises(ValueError):
        ckpt.resume_point(st, [1, 1, 5, 10])
    with pytest.raises(ValueError):
        ckpt.pack_state("warmup", 0, 0, 0, opt, sched, [1, 1, 5, 20])


def test_state_mismatch_and_missing_file(tmp_path):
    _, opt, sched = _make()
    st = ckpt.pack_state("disp_init", 0, 1, 1, opt, sched, [1, 1, 5, 20], with_rng=False)
    other = torch.optim.Adam(torch.nn.Linear(2, 2).parameters(), 1e-2)
    with pytest.raises(ValueError):
        ckpt.apply_state(st, other, torch.optim.lr_scheduler.StepLR(other, 2, 0.5))
    with pytest.raises(FileNotFoundError):
        ckpt.load_state(str(tmp_path))


def test_apply_state_refuses_a_different_parameter_order():
    """Adam moments are stored by parameter index; a checkpoint written with another parameter order must be
    refused instead of attaching the moments to the wrong tensors."""
    import pytest
    import torch
    from dd_b200 import checkpoint as ckpt

    a, b = torch.nn.Parameter(torch.ones(3)), torch.nn.Parameter(torch.ones(3))
    opt = torch.optim.Adam([a, b], 1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, 10, 0.5)
    (a.sum() + 2 * b.sum()).backward()
    opt.step()
    state = ckpt.pack_state("disp_init", 0, 1, 1, opt, sched, [1, 1, 1, 1], with_rng=False,
                            param_names=[("enc.a", (3,)), ("dec.b", (3,))])
    opt2 = torch.optim.Adam([b, a], 1e-3)
    sched2 = torch.optim.lr_scheduler.StepLR(opt2, 10, 0.5)
    with pytest.raises(ValueError, match="parameter order"):
        ckpt.apply_state(state, opt2, sched2, restore_rng=False, param_names=["dec.b", "enc.a"])
    ckpt.apply_state(state, opt2, sched2, restore_rng=False, param_names=["enc.a", "dec.b"])   # same order: accepted
