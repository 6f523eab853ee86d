"""Training-state checkpoint ("a real resume").

The reference writes the per-module weights and `adam.pth` (Trainer.py:697-707) but never reads the optimiser back and
keeps no record of where in the four-phase schedule (Trainer.py:466-490) a run stopped, so an interrupted run restarts
from `disp_init`.  Here one extra file, `trainer_state.pth`, sits next to them: phase, epoch / step counters, optimiser
and LR-scheduler state and the RNG streams; `Trainer.train(resume_from=folder)` continues with the next epoch of that
phase.  Pure host-side functions (no CUDA needed) so that they are covered by the CPU test tier.
"""
import os.path as osp

import numpy as np
import torch

PHASES = ["disp_init", "motion_init", "mask_init", "fine_tune"]
STATE_FILE = "trainer_state.pth"
FORMAT_VERSION = 1


def pack_state(phase_name, epoch, step, g_step, optimizer, lr_scheduler, epoch_schedules, with_rng=True, param_names=None):
    """Everything needed to continue after `epoch` (0-based, completed) of `phase_name`.
    param_names: [("<module>.<param>", shape), ...] in optimiser order -- Adam moments are stored by index, so the
    resuming process must attach them to the same parameters (checked by apply_state)."""
    if phase_name not in PHASES:
        raise ValueError(f"unknown phase {phase_name!r}")
    state = {
        "version": FORMAT_VERSION,
        "phase_name": phase_name, "epoch": int(epoch), "step": int(step), "g_step": int(g_step),
        "epoch_schedules": [int(e) for e in epoch_schedules],
        "optimizer": optimizer.state_dict(), "lr_scheduler": lr_scheduler.state_dict(),
    }
    if param_names is not None:
        state["param_names"] = [(str(n), tuple(int(d) for d in shp)) for n, shp in param_names]
    if with_rng:
        rng = {"torch_cpu": torch.get_rng_state(), "numpy": np.random.get_state()}
        if torch.cuda.is_available():
            rng["torch_cuda"] = torch.cuda.get_rng_state()
        state["rng"] = rng
    return state


def save_state(folder, state):
    torch.save(state, osp.join(folder, STATE_FILE))


def load_state(folder):
    path = osp.join(folder, STATE_FILE)
    if not osp.exists(path):
        raise FileNotFoundError(f"{path} not found: the folder holds weights only (written by the reference or by an older run); "
                                "use --load_ckpt to start a new schedule from those weights")
    state = torch.load(path, map_location="cpu", weights_only=False)
    if state.get("version") != FORMAT_VERSION:
        raise ValueError(f"{path}: unsupported trainer-state version {state.get('version')}")
    return state


def restore_rng_state(state):
    if "rng" in state:
        torch.set_rng_state(state["rng"]["torch_cpu"])
        np.random.set_state(state["rng"]["numpy"])
        if "torch_cuda" in state["rng"] and torch.cuda.is_available():
            torch.cuda.set_rng_state(state["rng"]["torch_cuda"])


def apply_state(state, optimizer, lr_scheduler, restore_rng=True, param_names=None):
    """Load optimiser / scheduler (and RNG) state; the caller has already built them for state['phase_name'].
    param_names: this process's optimiser-order parameter names; compared with the checkpoint's record."""
    if param_names is not None and "param_names" in state:
        saved_names = [n for n, _ in state["param_names"]]
        if list(param_names) != saved_names:
            diff = next((i for i, (a, b) in enumerate(zip(param_names, saved_names)) if a != b), min(len(param_names), len(saved_names)))
            raise ValueError(f"optimiser parameter order differs from the checkpoint's at index {diff}: "
                             f"{param_names[diff] if diff < len(param_names) else None!r} vs "
                             f"{saved_names[diff] if diff < len(saved_names) else None!r}")
    own, saved = optimizer.state_dict()["param_groups"], state["optimizer"]["param_groups"]
    if [len(g["params"]) for g in own] != [len(g["params"]) for g in saved]:
        raise ValueError("optimiser state does not match the phase's parameter list "
                         f"({[len(g['params']) for g in saved]} saved vs {[len(g['params']) for g in own]} now)")
    optimizer.load_state_dict(state["optimizer"])
    lr_scheduler.load_state_dict(state["lr_scheduler"])
    if restore_rng:
        restore_rng_state(state)


def resume_point(state, epoch_schedules):
    """(phase index, first epoch to run in that phase): the epoch after the saved one, or the next phase with epochs."""
    if [int(e) for e in epoch_schedules] != state["epoch_schedules"]:
        raise ValueError(f"epoch_schedules {list(epoch_schedules)} differ from the checkpoint's {state['epoch_schedules']}")
    phase_i, epoch = PHASES.index(state["phase_name"]), state["epoch"] + 1
    if epoch >= epoch_schedules[phase_i]:
        phase_i, epoch = phase_i + 1, 0
    return phase_i, epoch
