"""SURVEY 8f-3 opt-ins of Model.forward (reference networks/model.py:69-74 and :82-86): skipping the depth passes on
frames -1/+1 (no loss term reads them) and batching the two pose-encoder calls.  With BatchNorm / DropPath in eval mode
both are the same function as the default path, so `('disp', 0, s)`, the loss and every optimised-weight gradient must be
unchanged (they differ in train mode only through BatchNorm batch statistics / the DropPath RNG stream -- hence opt-in)."""
import pytest
import torch

from oracle.compare import check_rel

pytestmark = pytest.mark.gpu


def _trainer(extra=()):
    import options
    from Trainer import Trainer

    opt = options.DynamoOptions().parse(args=["-d", "waymo", "--depth_model", "litemono", "--weights_init", "scratch", "-b", "2", "--height", "64",
                                              "--width", "96", "--g_d_ground", "0.0", *extra])
    opt.ddp = False
    torch.manual_seed(7)
    tr = Trainer(opt)
    tr.setup_phase("fine_tune")
    tr.step, tr.num_steps_per_epoch = 100, 100
    tr.set_eval()
    return tr


def _step(tr, batch):
    tr.arena.zero()
    outputs, losses = tr.process_batch({k: v.cuda() for k, v in batch.items()})
    losses["loss"].backward()
    torch.cuda.synchronize()
    return outputs, losses, tr.arena.flat.clone()


@pytest.mark.parametrize("flag", ["--skip_unused_depth", "--batch_pose_pairs"])
def test_opt_in_forward_variants_leave_loss_and_gradients_unchanged(flag):
    from dd_b200 import synthetic

    torch.backends.cudnn.allow_tf32 = False
    base = _trainer()
    var = _trainer([flag])
    assert getattr(var.base_model, flag[2:]) is True and getattr(base.base_model, flag[2:]) is False
    var.base_model.load_state_dict(base.base_model.state_dict())
    batch = synthetic.make_batch(base.opt, 5)
    o0, l0, g0 = _step(base, batch)
    o1, l1, g1 = _step(var, batch)
    for s in base.opt.scales:
        check_rel(o1[("disp", 0, s)], o0[("disp", 0, s)], 1e-6, what=f"{flag} disp scale {s}")
    for f in (-1, 1):
        check_rel(o1[("cam_T_cam", 0, f)], o0[("cam_T_cam", 0, f)], 1e-5, what=f"{flag} cam_T_cam {f}")
        if flag == "--skip_unused_depth":
            assert ("disp", f, 0) in o0 and ("disp", f, 0) not in o1        # the dead passes are really gone
    check_rel(float(l1["loss"]), float(l0["loss"]), 1e-5, what=f"{flag} loss")
    assert base.arena.names == var.arena.names
    check_rel(g1, g0, 1e-4, what=f"{flag} all optimised-weight gradients (max-abs / max)")
    rel_l2 = float((g1 - g0).norm() / g0.norm())
    assert rel_l2 <= 1e-5, rel_l2
