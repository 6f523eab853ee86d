"""Host-side data-parallel logic (gradient arena + all-reduce) on CPU with the gloo backend, world size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dd_b200.parallel import GradArena

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    shared = list(net.parameters())
    arena = GradArena(shared + shared[:1], world_size=world)      # duplicate entry: shared encoder case
    assert arena.numel == sum(p.numel() for p in shared)
    x = torch.randn(5, 8, generator=torch.Generator().manual_seed(100 + rank))
    net(x).pow(2).mean().backward()
    assert arena.check_views(), "autograd must accumulate into the arena views in place"
    local = arena.flat.clone()
    arena.all_reduce()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = torch.stack(gathered).mean(0)
    ok = torch.allclose(arena.flat, expect, rtol=1e-6, atol=1e-8)
    # a second backward keeps accumulating into the same storage after zero()
    arena.zero()
    net(x).pow(2).mean().backward()
    ok = ok and arena.check_views() and torch.allclose(arena.flat, local, rtol=1e-6, atol=1e-8)
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_grad_arena_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(out.get(timeout=10) for _ in range(2))
    assert results == {0: True, 1: True}


def test_grad_arena_single_process_is_noop():
    from dd_b200.parallel import GradArena

    w = torch.nn.Parameter(torch.ones(3, 3))
    arena = GradArena([w], world_size=1)
    (w * 2).sum().backward()
    before = arena.flat.clone()
    arena.all_reduce()
    assert torch.equal(arena.flat, before) and torch.equal(w.grad, torch.full((3, 3), 2.0))


def test_library_exports_every_declared_symbol():
    """include/dynamo_b200.h <-> libdynamo_b200.so <-> ctypes signatures stay in sync (no compute calls)."""
    import re
    from dd_b200 import _lib

    lib = _lib.load()
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "dynamo_b200.h")).read()
    declared = set(re.findall(r"\b(dd_[a-z0-9_]+)\s*\(", header))
    assert declared, "no entry points found in the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dynamo_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.dd_version() >= 100
    assert set(_lib.SIGNATURES) <= declared


def test_cpu_tensors_are_refused_by_every_op():
    import tools
    from dd_b200 import DynamoB200Error
    from dd_b200 import functional as Fn

    x = torch.rand(1, 3, 8, 8)
    with pytest.raises(DynamoB200Error):
        tools.SSIM()(x, x)
    with pytest.raises(DynamoB200Error):
        Fn.conv2d_fused(x, torch.rand(4, 3, 3, 3))
    with pytest.raises(DynamoB200Error):
        tools.BackprojectDepth(1, 8, 8)(torch.rand(1, 1, 8, 8), torch.eye(4)[None])
    with pytest.raises(DynamoB200Error):
        tools.compute_smooth_loss(x, x)
    with pytest.raises(DynamoB200Error):
        Fn.linear(torch.rand(8, 8), torch.rand(4, 8), torch.rand(4))
    with pytest.raises(DynamoB200Error):
        Fn.nchw_to_nhwc(x)
    with pytest.raises(DynamoB200Error):
        Fn.block_tail(x, x.permute(0, 2, 3, 1).contiguous(), torch.rand(3))
    # round-2 encoder kernels: BatchNorm (+GELU), LayerNorm, depth-wise convolution, max-pool, XCA core
    bn = torch.nn.BatchNorm2d(4).train()
    x4 = torch.rand(2, 4, 8, 8)
    with pytest.raises(DynamoB200Error):
        Fn._BatchNormGeluFn.apply(x4, bn.weight, bn.bias, bn.running_mean, bn.running_var, 0.1, 1e-5, True)
    with pytest.raises(DynamoB200Error):
        Fn._BnActNHWCFn.apply(x4, None, bn.weight, bn.bias, bn.running_mean, bn.running_var, 0.1, 1e-5, 1)
    with pytest.raises(DynamoB200Error):
        Fn.layer_norm(torch.rand(4, 8), torch.ones(8), torch.zeros(8))
    with pytest.raises(DynamoB200Error):
        Fn.dwconv3x3(x4, torch.rand(4, 1, 3, 3), 1)
    with pytest.raises(DynamoB200Error):
        Fn.maxpool3x3s2(x4)
    with pytest.raises(DynamoB200Error):
        Fn.xca_core(torch.rand(1, 8, 96), torch.ones(4, 1, 1), 4)


def test_encoder_layers_keep_the_reference_formulation_on_cpu():
    """The oracle-side tests run the product's Lite-Mono modules on the CPU: there EncoderLinear is F.linear and the
    blocks use the reference's permute / layer-scale / DropPath / residual formulation (no kernel involved)."""
    from networks import depth_encoder as de

    torch.manual_seed(0)
    blk = de.DilatedConv(dim=8, k=3, dilation=1, drop_path=0.5).train()
    x = torch.rand(4, 8, 6, 10)
    torch.manual_seed(1)
    got = blk(x)
    torch.manual_seed(1)
    y = blk.bn1(blk.ddwconv(x)).permute(0, 2, 3, 1)
    y = blk.gamma * torch.nn.functional.linear(blk.act(torch.nn.functional.linear(y, blk.pwconv1.weight, blk.pwconv1.bias)),
                                                 blk.pwconv2.weight, blk.pwconv2.bias)
    mask = x.new_empty(4, 1, 1, 1).bernoulli_(0.5).div_(0.5)
    assert torch.allclose(got, x + y.permute(0, 3, 1, 2) * mask)


def test_options_defaults_match_reference_surface():
    import options

    opt = options.DynamoOptions().parse(args=[])
    assert (opt.dataset, opt.height, opt.width, opt.scales, opt.batch_size) == ("waymo", 320, 480, [0, 1, 2], 3)
    assert opt.frame_ids == [0, -1, 1] and opt.epoch_schedules == [1, 1, 5, 20] and opt.epoch_size == 8000
    opt = options.DynamoOptions().parse(args=["-d", "kitti", "--depth_model", "monodepthv2"])
    assert (opt.height, opt.width, opt.scales, opt.split) == (192, 640, [0, 1, 2, 3], "eigen_zhou")
    g = sorted(k for k in vars(opt) if k.startswith("g_"))
    assert g == ["g_c_consistency", "g_c_smooth", "g_d_ground", "g_d_smooth", "g_m_smooth", "g_m_sparsity", "g_p_photo"]


# ---------------------------------------------------------------------------------------------------------------
# round 2: replica synchronisation at construction, chunked / overlapped all-reduce, layout check, fixed module order
def _spawn(target, world=2, timeout=180, extra=()):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, out) + tuple(extra)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout)
        assert p.exitcode == 0
    return dict(out.get(timeout=10) for _ in range(world))


def _make_net(seed):
    torch.manual_seed(seed)
    return torch.nn.ModuleDict({
        "enc": torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.BatchNorm2d(4), torch.nn.ReLU()),
        "dec": torch.nn.Sequential(torch.nn.Conv2d(4, 2, 3, padding=1)),
        "head": torch.nn.Sequential(torch.nn.Linear(2, 3)),
    })


def _broadcast_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dd_b200.parallel import broadcast_module_state

    net = _make_net(100 + rank)                       # every rank starts from different weights, as with unseeded train.py
    net["enc"][1].running_mean.fill_(float(rank + 1))
    net["enc"][1].num_batches_tracked.fill_(7 * (rank + 1))
    n = broadcast_module_state(net, src=0)
    flat = torch.cat([t.detach().double().reshape(-1) for t in list(net.parameters()) + list(net.buffers())])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    ok = all(torch.equal(g, gathered[0]) for g in gathered) and n == flat.numel()
    ok = ok and float(net["enc"][1].running_mean[0]) == 1.0 and int(net["enc"][1].num_batches_tracked) == 7
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_construction_broadcast_makes_replicas_identical():
    assert _spawn(_broadcast_worker) == {0: True, 1: True}


def _overlap_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dd_b200.parallel import GradArena, broadcast_module_state

    net = _make_net(100 + rank)
    broadcast_module_state(net)
    named = [(f"{m}.{k}", p) for m in ("enc", "dec", "head") for k, p in net[m].named_parameters()]
    arena = GradArena([p for _, p in named], world_size=world, names=[n for n, _ in named],
                      chunk_ids=[n.split(".")[0] for n, _ in named], overlap=True)
    ok = len(arena.chunks) == 3 and arena.chunks[0][0] == 0 and arena.chunks[-1][1] == arena.numel
    opt = torch.optim.SGD([p for _, p in named], lr=0.1)
    for step in range(3):
        x = torch.randn(2, 3, 6, 6, generator=torch.Generator().manual_seed(10 * step + rank))
        # `head` receives no gradient at all: its chunk must still be reduced (zeros) and never block the others
        import copy
        twin = copy.deepcopy(net)                       # same weights, plain autograd: the local gradient before any exchange
        g = torch.autograd.grad(twin["dec"](twin["enc"](x)).pow(2).mean(), list(twin["enc"].parameters()) + list(twin["dec"].parameters()))
        local = torch.cat([t.reshape(-1) for t in g] + [torch.zeros(sum(p.numel() for p in net["head"].parameters()))])
        loss = net["dec"](net["enc"](x)).pow(2).mean()
        loss.backward()
        issued_in_backward = len(arena._works)
        arena.all_reduce()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        ok = ok and torch.allclose(arena.flat, torch.stack(gathered).mean(0), rtol=1e-6, atol=1e-9)
        ok = ok and arena.last_collectives == 3 and arena.check_views()
        # step 0 only records which parameters receive gradients; afterwards the chunks leave from the autograd hooks
        ok = ok and (issued_in_backward == 0 if step == 0 else issued_in_backward == 3)
        opt.step()
        arena.zero()
    flat = torch.cat([p.detach().reshape(-1) for _, p in named])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    ok = ok and torch.equal(gathered[0], gathered[1])          # replicas stay bit-identical
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_chunked_allreduce_from_autograd_hooks_gloo_world2():
    assert _spawn(_overlap_worker) == {0: True, 1: True}


def _layout_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dd_b200.parallel import GradArena

    net = _make_net(0)
    named = [(f"{m}.{k}", p) for m in (("enc", "dec") if rank == 0 else ("dec", "enc")) for k, p in net[m].named_parameters()]
    try:
        GradArena([p for _, p in named], world_size=world, names=[n for n, _ in named])
        ok = False
    except RuntimeError as e:
        ok = "layout" in str(e)
    out.put((rank, ok))
    dist.destroy_process_group()


def test_arena_layout_mismatch_between_ranks_is_refused():
    assert _spawn(_layout_worker) == {0: True, 1: True}


def test_module_and_parameter_order_do_not_depend_on_the_hash_seed():
    """list(set(...)) over module names made the arena / Adam-state order differ between processes."""
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r); import options, networks\n"
            "from dd_b200.parallel import layout_digest\n"
            "opt = options.DynamoOptions().parse(args=['--weights_init', 'scratch', '--height', '64', '--width', '96'])\n"
            "m = networks.Model(opt)\n"
            "named = m.named_parameters_by_names(['Depth', 'Pose', 'CmpFlow', 'MotMask'])\n"
            "assert [id(p) for _, p in named] == [id(p) for p in m.parameters_by_names(['MotMask', 'CmpFlow', 'Pose', 'Depth'])]\n"
            "print(','.join(m.module_names), layout_digest([(n, p.numel()) for n, p in named]))\n") % PKG_DIR
    outs = set()
    for seed in ("1", "2", "77"):
        env = dict(os.environ, PYTHONHASHSEED=seed)
        outs.add(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.strip())
    assert len(outs) == 1, outs
    assert outs.pop().startswith("depth_enc,depth_dec,pose_enc,pose_dec,motion_enc,motion_dec,motion_mask ")


PKG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dynamo-depth_b200")


def test_arena_views_follow_channels_last_parameters_and_detect_detached_grads():
    """R2 finding: ResNet trunks are converted to channels_last; Module.to(memory_format=...) re-creates .grad tensors, which
    silently detached them from the arena (their gradients were then never all-reduced).  The arena now mirrors the
    parameter's strides (the fused optimiser requires equal strides) and check_views() reports a detached gradient."""
    from dd_b200.parallel import GradArena

    conv = torch.nn.Conv2d(4, 6, 3).to(memory_format=torch.channels_last)
    lin = torch.nn.Linear(5, 3)
    params = list(conv.parameters()) + list(lin.parameters())
    arena = GradArena(params, world_size=1)
    assert conv.weight.grad.stride() == conv.weight.stride() != conv.weight.contiguous().stride()
    assert arena.check_views()
    x = torch.randn(2, 4, 8, 8).contiguous(memory_format=torch.channels_last)
    (conv(x).mean() + lin(torch.randn(2, 5)).sum()).backward()
    assert arena.check_views() and float(arena.flat.abs().sum()) > 0
    ref = torch.autograd.grad(torch.nn.functional.conv2d(x, conv.weight, conv.bias).mean(), conv.weight)[0]
    assert torch.allclose(conv.weight.grad, ref)
    # a late layout conversion replaces .grad: must be detected
    conv2 = torch.nn.Conv2d(4, 6, 3)
    arena2 = GradArena(list(conv2.parameters()), world_size=1)
    conv2.to(memory_format=torch.channels_last)
    assert not arena2.check_views()
