"""Data parallelism on real GPUs (NCCL, one process per GPU): what the reference gets from DistributedDataParallel
(Trainer.py:44,148) and SURVEY.md section 4 asks to test --

  * replicas are identical after construction although every rank seeded its weights differently,
  * the all-reduced gradient equals the gradient of ONE process run on the concatenated batch (BN / DropPath in eval
    mode so that the two are the same function), at 1e-5,
  * after three optimisation steps on different per-rank batches the weights of all ranks are bit-identical,
  * from the second step on the arena chunks leave from the autograd hooks (overlap with backward).

Skipped on boxes with fewer than two GPUs (`gpurun --gpus 2`).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

H, W, B_RANK = 96, 128, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _opt(batch, ddp, rank, log_dir):
    import options

    opt = options.DynamoOptions().parse(args=["-d", "waymo", "--depth_model", "litemono", "--weights_init", "scratch", "-b", str(batch),
                                              "--height", str(H), "--width", str(W), "--g_d_ground", "0.0", "--log_dir", log_dir])
    opt.ddp = ddp
    opt.local_rank = 0
    opt.cuda_ids = [rank]
    return opt


def _params_flat(tr, nets):
    return torch.cat([p.detach().reshape(-1) for p in tr.base_model.parameters_by_names(nets)])


def _same_on_all_ranks(t):
    hi, lo = t.clone(), t.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return bool(torch.equal(hi, lo))


def _worker(rank, world, port, out, log_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "dynamo-depth_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from Trainer import Trainer
    from dd_b200 import synthetic

    res = {}
    all_nets = ["Depth", "Pose", "CmpFlow", "MotMask"]
    torch.manual_seed(100 + rank)                          # different initial weights per rank, as with an unseeded train.py
    tr = Trainer(_opt(B_RANK, True, rank, log_dir))
    res["identical_after_construction"] = _same_on_all_ranks(_params_flat(tr, all_nets))
    bufs = torch.cat([b.detach().double().reshape(-1) for b in tr.base_model.buffers()])
    res["buffers_identical_after_construction"] = _same_on_all_ranks(bufs)
    snapshot = {k: v.detach().clone() for k, v in tr.base_model.state_dict().items()}

    # ---- gradient equality vs one process on the concatenated batch (disp_init without automask: every loss term is a
    # mean over the batch, so the mean of the per-rank gradients is the gradient of the big batch)
    batches = [synthetic.make_batch(tr.opt, 1000 + r, batch=B_RANK) for r in range(world)]
    tr.setup_phase("disp_init")
    tr.bool_automask = False
    tr.num_steps_per_epoch, tr.step = 10, 10
    tr.set_eval()
    _, losses = tr.process_batch({k: v.to(tr.device) for k, v in batches[rank].items()})
    losses["loss"].backward()
    tr.arena.all_reduce()
    g_ddp = tr.arena.flat.clone()
    res["chunks"] = len(tr.arena.chunks)
    tr.arena.zero()
    if rank == 0:
        one = Trainer(_opt(B_RANK * world, False, rank, log_dir))
        one.base_model.load_state_dict(snapshot)
        one.setup_phase("disp_init")
        one.bool_automask = False
        one.num_steps_per_epoch, one.step = 10, 10
        one.set_eval()
        cat = {k: torch.cat([b[k] for b in batches], 0).to(one.device) for k in batches[0]}
        _, l1 = one.process_batch(cat)
        l1["loss"].backward()
        g_one = one.arena.flat
        assert one.arena.names == tr.arena.names
        res["grad_rel_l2"] = float((g_ddp - g_one).norm() / g_one.norm())
        res["grad_max_abs_rel"] = float((g_ddp - g_one).abs().max() / g_one.abs().max())
        worst = 0.0
        off = 0
        for n, p in zip(one.arena.names, one.arena.params):       # per tensor, relative to the largest gradient tensor norm
            d = (g_ddp[off:off + p.numel()] - g_one[off:off + p.numel()]).norm()
            worst = max(worst, float(d / g_one.norm()))
            off += p.numel()
        res["grad_worst_tensor"] = worst
        del one
    dist.barrier()

    # ---- three optimisation steps in train mode (fine_tune: all four networks, hooks issue the chunks during backward)
    tr.setup_phase("fine_tune")
    tr.num_steps_per_epoch, tr.step = 10, 10
    tr.set_train()
    issued = []
    before = _params_flat(tr, all_nets).clone()
    for i in range(3):
        b = synthetic.make_batch(tr.opt, 2000 + 10 * i + rank, batch=B_RANK)
        orig = tr.arena.all_reduce

        def spy(*a, _orig=orig, **kw):
            issued.append(len(tr.arena._works))
            return _orig(*a, **kw)
        tr.arena.all_reduce = spy
        tr.train_step({k: v.to(tr.device) for k, v in b.items()})
        tr.arena.all_reduce = orig
    res["issued_from_hooks"] = issued
    res["collectives_per_step"] = tr.arena.last_collectives
    res["identical_after_3_steps"] = _same_on_all_ranks(_params_flat(tr, all_nets))
    res["weights_moved"] = float((_params_flat(tr, all_nets) - before).abs().max()) > 0
    torch.cuda.synchronize()
    out.put((rank, res))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_ddp_replicas_and_gradients_nccl_world2(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    import queue
    import time
    results, t0 = {}, time.time()
    while len(results) < world and time.time() - t0 < 900:
        try:
            r, res = out.get(timeout=5)
            results[r] = res
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead:                       # a rank died: do not wait for the NCCL timeout of the others
                for p in procs:
                    if p.is_alive():
                        p.kill()
                pytest.fail(f"a rank exited with {dead}")
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert len(results) == world
    for r in range(world):
        res = results[r]
        assert res["identical_after_construction"] and res["buffers_identical_after_construction"], res
        assert res["identical_after_3_steps"] and res["weights_moved"], res
        assert res["chunks"] == 4 and res["collectives_per_step"] == 7, res          # disp_init: 4 sub-modules; fine_tune: 7
        assert res["issued_from_hooks"][0] == 0 and all(n >= 5 for n in res["issued_from_hooks"][1:]), res
    r0 = results[0]
    assert r0["grad_rel_l2"] <= 1e-5 and r0["grad_worst_tensor"] <= 1e-5, r0
    try:
        from oracle import parity_log
        parity_log.record("ddp_nccl_world2", "arena_grad_vs_single_process", rel_l2=r0["grad_rel_l2"], max_abs_rel=r0["grad_max_abs_rel"])
    except Exception:
        pass
