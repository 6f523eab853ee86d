"""Trainer.train(resume_from=...): an interrupted four-phase schedule continues with the next epoch of the
phase the checkpoint was written in -- weights, Adam state, LR schedule and counters restored."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _opt(tmp_path, name, extra=()):
    import options
    opt = options.DynamoOptions().parse(args=["-d", "kitti", "--depth_model", "litemono", "--weights_init", "scratch", "-b", "2",
                                              "--height", "64", "--width", "96", "--g_d_ground", "0.0", "--epoch_schedules", "2", "1", "0",
                                              "0", "--log_dir", str(tmp_path), "-n", name, *extra])
    opt.ddp = False
    opt.epoch_size = 2
    return opt


def test_trainer_resume_continues_schedule(tmp_path):
    from Trainer import Trainer
    from dd_b200 import checkpoint as ckpt

    tr = Trainer(_opt(tmp_path, "full"))
    tr.train()
    models = os.path.join(str(tmp_path), "full", "models")
    folders = lambda d: sorted(f for f in os.listdir(d) if os.path.isdir(os.path.join(d, f)))
    assert os.path.isfile(os.path.join(models, "opt.json"))          # written at construction, as the reference does (Trainer.py:85)
    assert folders(models) == ["disp_init_00", "disp_init_01", "motion_init_00"]
    first = ckpt.load_state(os.path.join(models, "disp_init_00"))
    assert (first["phase_name"], first["epoch"], first["step"], first["g_step"]) == ("disp_init", 0, 2, 2)
    assert tr.g_step == 6 and tr.phase_name == "motion_init"

    tr2 = Trainer(_opt(tmp_path, "resumed", ["--resume", os.path.join(models, "disp_init_00")]))
    w0 = torch.load(os.path.join(models, "disp_init_00", "depth_dec.pth"), map_location="cpu")
    tr2.train()
    # the run picked up the weights of the checkpoint and went on with disp_init epoch 1, then motion_init
    models2 = os.path.join(str(tmp_path), "resumed", "models")
    assert folders(models2) == ["disp_init_01", "motion_init_00"]
    assert tr2.g_step == 6 and tr2.phase_name == "motion_init"
    mid = ckpt.load_state(os.path.join(models2, "disp_init_01"))
    assert (mid["phase_name"], mid["epoch"], mid["step"], mid["g_step"]) == ("disp_init", 1, 4, 4)
    steps = {int(v["step"]) for v in mid["optimizer"]["state"].values()}
    assert steps == {4}, steps                                   # Adam's own step count continued from the checkpoint (2 + 2)
    ref = ckpt.load_state(os.path.join(models, "disp_init_01"))
    assert mid["lr_scheduler"]["last_epoch"] == ref["lr_scheduler"]["last_epoch"] == 2
    # same data (seeded synthetic triplets), same restored state: the resumed epoch lands close to the uninterrupted one
    a = torch.load(os.path.join(models, "disp_init_01", "depth_dec.pth"), map_location="cpu")
    b = torch.load(os.path.join(models2, "disp_init_01", "depth_dec.pth"), map_location="cpu")
    # (mean over all weights: Adam's sign-like update can flip on an isolated weight whose gradient is rounding noise)
    n = sum(a[k].numel() for k in a)
    moved = sum(float((a[k] - w0[k]).abs().sum()) for k in a) / n
    apart = sum(float((a[k] - b[k]).abs().sum()) for k in a) / n
    assert moved > 0 and apart <= 0.1 * moved, (moved, apart)
