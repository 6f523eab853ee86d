"""torch.autograd bindings of the fused sm_100a kernels (host side of the C ABI).

`view_synthesis_sums` is what Trainer.generate_images_pred / compute_losses dispatch to: one forward
launch covers Trainer.py:215-287 and the photometric / automask / c_consistency / disp_mag parts of
Trainer.py:312-397 for every pyramid level and both source frames; one backward launch produces the
gradients w.r.t. disp_s, cam_T_cam, complete_flow_s and motion_mask_s.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import _lib as L


@dataclass
class WarpConfig:
    scales: List[int]
    cmpflow: bool = False
    motmask: bool = False
    automask: bool = False
    min_depth: float = 0.1
    max_depth: float = 100.0
    ssim_weight: float = 0.85
    mask_disp_thrd: float = 0.03
    # names of by-products to materialise: 'warped', 'sample', 'depth', 'ident_sel', 'resid', 'independ', 'mag'
    materialise: tuple = ()
    keep_warped: bool = True   # keep ('color', f, s) for the backward pass (24 B per pixel and level) instead of re-warping
    aux: Dict = field(default_factory=dict)  # filled by the forward pass: (name, frame_index, level) -> tensor

    @property
    def flags(self):
        return (L.DD_FLAG_CMPFLOW if self.cmpflow else 0) | (L.DD_FLAG_MOTMASK if self.motmask else 0) | \
               (L.DD_FLAG_AUTOMASK if self.automask else 0)


# optional CUDA-event timers around individual C-ABI calls (bench.py's roofline leg):
# KERNEL_TIMERS[name] = list of (start_event, end_event) while enabled, None otherwise
KERNEL_TIMERS = None


class _timed:
    """CUDA events around one C-ABI call while KERNEL_TIMERS is a dict.  `detail` timers (the many linear-layer calls of a
    step) are only taken when KERNEL_TIMERS["__detail__"] is set, and carry `meta` (e.g. flops / bytes of the call)."""

    def __init__(self, name, meta=None, detail=False):
        self.name, self.meta, self.detail = name, meta, detail
        self.ev = None

    def __enter__(self):
        if KERNEL_TIMERS is not None and (not self.detail or KERNEL_TIMERS.get("__detail__")):
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None:
            self.ev[1].record()
            KERNEL_TIMERS.setdefault(self.name, []).append(self.ev if self.meta is None else self.ev + (self.meta,))
        return False


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _prep(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _fill_desc(cfg, target, sources, K, inv_K, Ts, tss, noises, disps, flows, masks):
    d = L.WarpDesc()
    B, _, H, W = target.shape
    d.B, d.H, d.W = B, H, W
    d.num_scales = len(cfg.scales)
    d.num_frames = len(sources)
    d.flags = cfg.flags
    d.min_depth, d.max_depth = cfg.min_depth, cfg.max_depth
    d.ssim_weight, d.mask_disp_thrd = cfg.ssim_weight, cfg.mask_disp_thrd
    d.target = L.ptr(target)
    d.K, d.inv_K = L.ptr(K), L.ptr(inv_K)
    for f, src in enumerate(sources):
        assert src.shape == target.shape
        d.source[f] = L.ptr(src)
        assert Ts[f].shape == (B, 4, 4)
        d.T[f] = L.ptr(Ts[f])
        d.ts[f] = L.ptr(tss[f]) if tss[f] is not None else None
    for i, s in enumerate(cfg.scales):
        h, w = H >> s, W >> s
        d.scale[i] = s
        assert disps[i].shape == (B, 1, h, w), (disps[i].shape, (B, 1, h, w))
        d.disp[i] = L.ptr(disps[i])
        d.noise[i] = L.ptr(noises[i]) if noises and noises[i] is not None else None
        for f in range(len(sources)):
            if cfg.cmpflow:
                assert flows[i][f].shape == (B, 3, h, w)
                d.flow[i][f] = L.ptr(flows[i][f])
            if cfg.motmask:
                assert masks[i][f].shape == (B, 1, h, w)
                d.mask[i][f] = L.ptr(masks[i][f])
    return d


class _ViewSynthesisFn(torch.autograd.Function):
    """inputs: cfg, nF, target, K, inv_K, [src]*F, [ts]*F, [noise]*S, [T]*F, then per level: disp, [flow]*F, [mask]*F"""

    @staticmethod
    def forward(ctx, cfg, nF, *tensors):
        lib = L.load()
        S = len(cfg.scales)
        it = iter(tensors)
        target, K, inv_K = next(it), next(it), next(it)
        sources = [next(it) for _ in range(nF)]
        tss = [next(it) for _ in range(nF)]
        noises = [next(it) for _ in range(S)]
        Ts = [next(it) for _ in range(nF)]
        disps, flows, masks = [], [], []
        for _ in range(S):
            disps.append(next(it))
            flows.append([next(it) for _ in range(nF)] if cfg.cmpflow else None)
            masks.append([next(it) for _ in range(nF)] if cfg.motmask else None)
        B, _, H, W = target.shape
        dev = target.device
        desc = _fill_desc(cfg, target, sources, K, inv_K, Ts, tss, noises, disps, flows, masks)

        want = set(cfg.materialise)
        if cfg.motmask:
            want |= {"resid", "mag"}   # resid: needed by backward (levels > 0); mag: m_sparsity
        if cfg.keep_warped and any(ctx.needs_input_grad):
            want |= {"warped"}         # backward re-reads the warped frames instead of re-warping them
        aux = L.WarpAux()
        out = {}
        for i, s in enumerate(cfg.scales):
            h, w = H >> s, W >> s
            if "depth" in want:
                out[("depth", 0, i)] = torch.empty(B, 1, H, W, device=dev)
                aux.depth[i] = out[("depth", 0, i)].data_ptr()
            if "ident_sel" in want and cfg.automask:
                out[("ident_sel", 0, i)] = torch.empty(B, H, W, device=dev)
                aux.ident_sel[i] = out[("ident_sel", 0, i)].data_ptr()
            for f in range(nF):
                if "warped" in want:
                    out[("warped", f, i)] = torch.empty(B, 3, H, W, device=dev)
                    aux.warped[i][f] = out[("warped", f, i)].data_ptr()
                if "sample" in want:
                    out[("sample", f, i)] = torch.empty(B, H, W, 2, device=dev)
                    aux.sample[i][f] = out[("sample", f, i)].data_ptr()
                if cfg.cmpflow and "resid" in want:
                    out[("resid", f, i)] = torch.empty(B, 3, h, w, device=dev)
                    aux.resid[i][f] = out[("resid", f, i)].data_ptr()
                if cfg.cmpflow and "independ" in want:
                    out[("independ", f, i)] = torch.empty(B, 3, H, W, device=dev)
                    aux.independ[i][f] = out[("independ", f, i)].data_ptr()
                if cfg.cmpflow and "mag" in want:
                    out[("mag", f, i)] = torch.empty(B, h, w, device=dev)
                    aux.mag[i][f] = out[("mag", f, i)].data_ptr()
        cfg.aux = out

        sums = torch.empty(S, L.DD_NSUM, device=dev)
        ws_bytes = lib.dd_warp_photo_workspace_bytes(C.byref(desc))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with _timed("warp_photo_fwd"):
            L.check(lib.dd_warp_photo_fwd(C.byref(desc), C.byref(aux), sums.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
                    "dd_warp_photo_fwd")

        ctx.cfg, ctx.nF = cfg, nF
        ctx.save_for_backward(*tensors)
        ctx.saved_aux = {k: v for k, v in out.items() if k[0] in ("resid", "warped")}
        return sums

    @staticmethod
    def backward(ctx, grad_sums):
        lib = L.load()
        cfg, nF = ctx.cfg, ctx.nF
        S = len(cfg.scales)
        tensors = ctx.saved_tensors
        it = iter(tensors)
        target, K, inv_K = next(it), next(it), next(it)
        sources = [next(it) for _ in range(nF)]
        tss = [next(it) for _ in range(nF)]
        noises = [next(it) for _ in range(S)]
        Ts = [next(it) for _ in range(nF)]
        disps, flows, masks = [], [], []
        for _ in range(S):
            disps.append(next(it))
            flows.append([next(it) for _ in range(nF)] if cfg.cmpflow else None)
            masks.append([next(it) for _ in range(nF)] if cfg.motmask else None)
        B, _, H, W = target.shape
        dev = target.device
        desc = _fill_desc(cfg, target, sources, K, inv_K, Ts, tss, noises, disps, flows, masks)

        # position of each differentiable input inside `tensors` (after cfg, nF)
        n_fixed = 3 + nF + nF + S
        need = ctx.needs_input_grad[2:]
        grads = [None] * len(tensors)
        g = L.WarpGrads()
        pos = n_fixed
        for f in range(nF):
            if need[pos]:
                grads[pos] = torch.empty(B, 4, 4, device=dev)
                g.T[f] = grads[pos].data_ptr()
            pos += 1
        any_T = any(grads[n_fixed + f] is not None for f in range(nF))
        if any_T:   # the finalize kernel writes both frames' matrices
            for f in range(nF):
                if grads[n_fixed + f] is None:
                    grads[n_fixed + f] = torch.empty(B, 4, 4, device=dev)
                    g.T[f] = grads[n_fixed + f].data_ptr()
        for i in range(S):
            if need[pos]:
                grads[pos] = torch.empty_like(disps[i])
                g.disp[i] = grads[pos].data_ptr()
            pos += 1
            if cfg.cmpflow:
                for f in range(nF):
                    if need[pos]:
                        grads[pos] = torch.empty_like(flows[i][f])
                        g.flow[i][f] = grads[pos].data_ptr()
                    pos += 1
            if cfg.motmask:
                for f in range(nF):
                    if need[pos]:
                        grads[pos] = torch.empty_like(masks[i][f])
                        g.mask[i][f] = grads[pos].data_ptr()
                    pos += 1

        saved = L.WarpAux()
        for (name, f, i), t in ctx.saved_aux.items():
            if name == "resid":
                saved.resid[i][f] = t.data_ptr()
            else:
                saved.warped[i][f] = t.data_ptr()
        gs = grad_sums.contiguous().float()
        ws_bytes = lib.dd_warp_photo_workspace_bytes(C.byref(desc))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with _timed("warp_photo_bwd"):
            L.check(lib.dd_warp_photo_bwd(C.byref(desc), gs.data_ptr(), C.byref(saved), C.byref(g), ws.data_ptr(), ws_bytes,
                                          _stream()), "dd_warp_photo_bwd")
        for idx in range(len(grads)):
            if not need[idx]:
                grads[idx] = None
        return (None, None) + tuple(grads)


def view_synthesis_sums(cfg: WarpConfig, target, sources, K, inv_K, Ts, tss, disps, flows=None, masks=None, noises=None):
    """Returns sums (S, DD_NSUM) (see include/dynamo_b200.h) -- differentiable w.r.t. Ts, disps, flows, masks.
    By-products requested through cfg.materialise (and those the scene-flow phases always need) land in cfg.aux."""
    nF = len(sources)
    S = len(cfg.scales)
    if not target.is_cuda:
        raise L.DynamoB200Error("view_synthesis_sums needs CUDA tensors (no CPU fallback)")
    B = target.shape[0]
    args = [_prep(target), _prep(K), _prep(inv_K)]
    args += [_prep(s) for s in sources]
    args += [(_prep(t).reshape(B) if t is not None else None) for t in (tss or [None] * nF)]
    args += [(_prep(n) if n is not None else None) for n in (noises or [None] * S)]
    args += [_prep(T) for T in Ts]
    for i in range(S):
        args.append(_prep(disps[i]))
        if cfg.cmpflow:
            args += [_prep(x) for x in flows[i]]
        if cfg.motmask:
            args += [_prep(x) for x in masks[i]]
    return _ViewSynthesisFn.apply(cfg, nF, *args)


# ---------------------------------------------------------------------------------------------
# edge-aware smoothness (tools.compute_smooth_loss, tools.py:311-326), batched
# ---------------------------------------------------------------------------------------------


def _smooth_tasks(inps, imgs, norms, grads=None):
    n = len(inps)
    arr = (L.SmoothTask * n)()
    for i in range(n):
        B, Cc, h, w = inps[i].shape
        arr[i].inp = L.ptr(inps[i])
        arr[i].img = L.ptr(imgs[i]) if imgs[i] is not None else None
        arr[i].grad_inp = grads[i].data_ptr() if grads is not None and grads[i] is not None else None
        arr[i].B, arr[i].C, arr[i].h, arr[i].w = B, Cc, h, w
        arr[i].mean_normalise = 1 if norms[i] else 0
        if imgs[i] is not None:
            assert imgs[i].shape == (B, 3, h, w), (imgs[i].shape, inps[i].shape)
    return arr


class _SmoothFn(torch.autograd.Function):
    """inputs: norms (tuple of bool), n, inp_0..inp_{n-1}, img_0..img_{n-1}; returns (n, 2) sums."""

    @staticmethod
    def forward(ctx, norms, n, *tensors):
        lib = L.load()
        inps, imgs = tensors[:n], tensors[n:]
        tasks = _smooth_tasks(inps, imgs, norms)
        dev = inps[0].device
        sums = torch.empty(n, 2, device=dev)
        nbytes = lib.dd_smooth_workspace_bytes(tasks, n)
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        L.check(lib.dd_smooth_fwd(tasks, n, sums.data_ptr(), ws.data_ptr(), nbytes, _stream()), "dd_smooth_fwd")
        ctx.norms, ctx.n = norms, n
        ctx.save_for_backward(*tensors)
        return sums

    @staticmethod
    def backward(ctx, grad_sums):
        lib = L.load()
        n, norms = ctx.n, ctx.norms
        tensors = ctx.saved_tensors
        inps, imgs = tensors[:n], tensors[n:]
        need = ctx.needs_input_grad[2:2 + n]
        grads = [torch.empty_like(inps[i]) if need[i] else None for i in range(n)]
        tasks = _smooth_tasks(inps, imgs, norms, grads)
        dev = inps[0].device
        nbytes = lib.dd_smooth_workspace_bytes(tasks, n)
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        gs = grad_sums.contiguous().float()
        L.check(lib.dd_smooth_bwd(tasks, n, gs.data_ptr(), ws.data_ptr(), nbytes, _stream()), "dd_smooth_bwd")
        return (None, None) + tuple(grads) + (None,) * n


def smooth_sums(inps, imgs, norms):
    """inps[i] (B,C,h,w), imgs[i] (B,3,h,w) or None, norms[i] bool -> (n,2) tensor of (sum_x, sum_y)."""
    n = len(inps)
    if n == 0:
        raise ValueError("smooth_sums: no tasks")
    if not inps[0].is_cuda:
        raise L.DynamoB200Error("smooth_sums needs CUDA tensors (no CPU fallback)")
    return _SmoothFn.apply(tuple(bool(x) for x in norms), n, *[_prep(x) for x in inps], *[_prep(x) for x in imgs])


_SMOOTH_DEN = {}


def smooth_means(sums, shapes):
    """(n,2) sums -> per-task  mean_x + mean_y  (tools.py:326) as an (n,) tensor."""
    key = (tuple(tuple(int(v) for v in sh) for sh in shapes), str(sums.device))
    den = _SMOOTH_DEN.get(key)
    if den is None:   # built once per shape set: torch.tensor(list, device=cuda) is a pageable copy that blocks the host
        den = _SMOOTH_DEN[key] = torch.tensor([[B * Cc * h * (w - 1), B * Cc * (h - 1) * w] for (B, Cc, h, w) in shapes],
                                              dtype=torch.float32).to(sums.device)
    return (sums / den).sum(1)


# ---------------------------------------------------------------------------------------------
# motion-mask sparsity (Trainer.py:388-399)
# ---------------------------------------------------------------------------------------------


class _MSparsityFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mag, mag_sum, prob):
        lib = L.load()
        B, _, h, w = prob.shape
        dev = prob.device
        out = torch.empty(4, device=dev)
        nbytes = lib.dd_msparsity_workspace_bytes(B, h, w)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        L.check(lib.dd_msparsity_fwd(L.ptr(mag), mag_sum.data_ptr(), L.ptr(prob), B, h, w, out.data_ptr(), ws.data_ptr(),
                                     nbytes, _stream()), "dd_msparsity_fwd")
        ctx.save_for_backward(mag, mag_sum, prob, out)
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        lib = L.load()
        mag, mag_sum, prob, out = ctx.saved_tensors
        B, _, h, w = prob.shape
        grad_prob = torch.empty_like(prob)
        go = grad_out.contiguous().float().reshape(1)
        L.check(lib.dd_msparsity_bwd(L.ptr(mag), mag_sum.data_ptr(), L.ptr(prob), out.data_ptr(), go.data_ptr(), B, h, w,
                                     grad_prob.data_ptr(), _stream()), "dd_msparsity_bwd")
        return None, None, grad_prob


def motion_sparsity(mag, mag_sum, prob):
    """mag (B,h,w) and its batch sum (0-dim/1-elem device tensor) from the fused forward; prob (B,1,h,w)."""
    if not prob.is_cuda:
        raise L.DynamoB200Error("motion_sparsity needs CUDA tensors (no CPU fallback)")
    return _MSparsityFn.apply(_prep(mag), mag_sum.contiguous(), _prep(prob))


# ---------------------------------------------------------------------------------------------
# fused decoder convolution (layers.py:85-121 ConvBlock / Conv3x3 / upsample + skip concat)
# ---------------------------------------------------------------------------------------------

_WS_CACHE = {}


def _workspace(nbytes, dev):
    """Per-device scratch buffer reused across calls (stream-ordered, so reuse on one stream is safe)."""
    if dev.type != "cuda":
        raise L.DynamoB200Error("dynamo_b200 ops need CUDA tensors (no CPU fallback)")
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, device=dev)
        _WS_CACHE[key] = buf
    return buf


def _conv_desc(x0, x1, weight, bias, residual, ksize, pad_mode, act, up0):
    d = L.ConvDesc()
    B, C0, H0, W0 = x0.shape
    H, W = (H0, W0) if up0 == L.UP_NONE else (2 * H0, 2 * W0)
    Cout, Cin = weight.shape[0], weight.shape[1]
    C1 = x1.shape[1] if x1 is not None else 0
    if Cin != C0 + C1 or weight.shape[2] != ksize or weight.shape[3] != ksize:
        raise ValueError(f"conv weight {tuple(weight.shape)} does not match inputs C0={C0} C1={C1} k={ksize}")
    if x1 is not None and tuple(x1.shape) != (B, C1, H, W):
        raise ValueError(f"skip tensor {tuple(x1.shape)} != {(B, C1, H, W)}")
    d.B, d.H, d.W, d.Cout = B, H, W, Cout
    d.ksize, d.pad_mode, d.act, d.up0 = ksize, pad_mode, act, up0
    d.C0, d.C1 = C0, C1
    d.x0, d.x1 = L.ptr(x0), L.ptr(x1)
    d.weight, d.bias, d.residual = L.ptr(weight), L.ptr(bias), L.ptr(residual)
    return d, (B, Cout, H, W)


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, x1, weight, bias, residual, ksize, pad_mode, act, up0):
        lib = L.load()
        d, oshape = _conv_desc(x0, x1, weight, bias, residual, ksize, pad_mode, act, up0)
        out = torch.empty(oshape, device=x0.device)
        nbytes = lib.dd_conv_workspace_bytes(C.byref(d))
        ws = _workspace(nbytes, x0.device)
        L.check(lib.dd_conv_fwd(C.byref(d), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "dd_conv_fwd")
        ctx.cfg = (ksize, pad_mode, act, up0)
        ctx.has = (x1 is not None, bias is not None, residual is not None)
        ctx.save_for_backward(x0, x1, weight, bias, out if act != L.ACT_NONE else None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = L.load()
        x0, x1, weight, bias, out = ctx.saved_tensors
        ksize, pad_mode, act, up0 = ctx.cfg
        grad_out = grad_out.contiguous()
        d, _ = _conv_desc(x0, x1, weight, bias, None, ksize, pad_mode, act, up0)
        need = ctx.needs_input_grad
        gx0 = torch.empty_like(x0) if need[0] else None
        gx1 = torch.empty_like(x1) if (x1 is not None and need[1]) else None
        gw = torch.empty_like(weight) if need[2] else None
        gb = torch.empty_like(bias) if (bias is not None and need[3]) else None
        nbytes = lib.dd_conv_workspace_bytes(C.byref(d))
        ws = _workspace(nbytes, x0.device)
        p = lambda t: t.data_ptr() if t is not None else None
        L.check(lib.dd_conv_bwd(C.byref(d), p(out), grad_out.data_ptr(), p(gx0), p(gx1), p(gw), p(gb), ws.data_ptr(),
                                ws.numel(), _stream()), "dd_conv_bwd")
        gres = grad_out if (ctx.has[2] and need[4]) else None
        return gx0, gx1, gw, gb, gres, None, None, None, None


def conv2d_fused(x0, weight, bias=None, *, x1=None, residual=None, ksize=3, pad="reflect", act="none", up="none"):
    """out = act(conv_k(pad(cat(up(x0), x1))) + bias) [+ residual]   (hand-written sm_100a kernels, no fallback)."""
    if not x0.is_cuda:
        raise L.DynamoB200Error("conv2d_fused needs CUDA tensors (no CPU fallback)")
    pad_mode = {"zero": L.PAD_ZERO, "reflect": L.PAD_REFLECT}[pad]
    act_id = {"none": L.ACT_NONE, "elu": L.ACT_ELU, "sigmoid": L.ACT_SIGMOID, "relu": L.ACT_RELU}[act]
    up_id = {"none": L.UP_NONE, "nearest": L.UP_NEAREST2, "bilinear": L.UP_BILINEAR2}[up]
    return _ConvFn.apply(_prep(x0), _prep(x1), _prep(weight), _prep(bias), _prep(residual), ksize, pad_mode, act_id, up_id)


class _ResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size, sigmoid):
        lib = L.load()
        lead = x.shape[:-2]
        hi, wi = x.shape[-2:]
        ho, wo = size
        BC = int(x.numel() // (hi * wi))
        out = torch.empty(*lead, ho, wo, device=x.device)
        L.check(lib.dd_resize_bilinear_fwd(x.data_ptr(), BC, hi, wi, ho, wo, int(sigmoid), out.data_ptr(), _stream()),
                "dd_resize_bilinear_fwd")
        ctx.dims = (BC, hi, wi, ho, wo, int(sigmoid))
        ctx.save_for_backward(out if sigmoid else None)
        ctx.in_shape = x.shape
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = L.load()
        (out,) = ctx.saved_tensors
        BC, hi, wi, ho, wo, sig = ctx.dims
        grad_out = grad_out.contiguous()
        gx = torch.empty(ctx.in_shape, device=grad_out.device)
        L.check(lib.dd_resize_bilinear_bwd(grad_out.data_ptr(), out.data_ptr() if out is not None else None, BC, hi, wi, ho,
                                           wo, sig, gx.data_ptr(), _stream()), "dd_resize_bilinear_bwd")
        return gx, None, None


def resize_bilinear(x, size, sigmoid=False):
    """F.interpolate(x, size, mode='bilinear', align_corners=False) [+ sigmoid]  (utils.py:98-101)."""
    if not x.is_cuda:
        raise L.DynamoB200Error("resize_bilinear needs CUDA tensors (no CPU fallback)")
    return _ResizeFn.apply(_prep(x), (int(size[0]), int(size[1])), bool(sigmoid))


# ---------------------------------------------------------------------------------------------
# stand-alone geometry / SSIM layers (tools.py:167-257) -- module surface for eval / user code
# ---------------------------------------------------------------------------------------------


def pyramid_half(x):
    """clamp(bicubic-antialias x1/2 of x, 0, 1): one link of the input colour pyramid (Trainer.py:80, 729-734).
    x: (B, C, H, W) CUDA fp32 with even H, W; input data, no gradient."""
    if not x.is_cuda:
        raise L.DynamoB200Error("pyramid_half needs CUDA tensors (no CPU fallback)")
    x = _prep(x.detach())
    B, Cc, H, W = x.shape
    out = torch.empty(B, Cc, H // 2, W // 2, device=x.device)
    L.check(L.load().dd_pyramid_half_fwd(x.data_ptr(), B * Cc, H, W, out.data_ptr(), _stream()), "dd_pyramid_half_fwd")
    return out


class _BackprojectFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, inv_K):
        lib = L.load()
        B, _, H, W = depth.shape
        pts = torch.empty(B, 4, H * W, device=depth.device)
        L.check(lib.dd_backproject_fwd(L.ptr(depth), L.ptr(inv_K), B, H, W, pts.data_ptr(), _stream()), "dd_backproject_fwd")
        ctx.save_for_backward(inv_K)
        ctx.dims = (B, H, W)
        return pts

    @staticmethod
    def backward(ctx, gpts):
        lib = L.load()
        (inv_K,) = ctx.saved_tensors
        B, H, W = ctx.dims
        gd = torch.empty(B, 1, H, W, device=gpts.device)
        L.check(lib.dd_backproject_bwd(gpts.contiguous().data_ptr(), inv_K.data_ptr(), B, H, W, gd.data_ptr(), _stream()),
                "dd_backproject_bwd")
        return gd, None


def backproject(depth, inv_K):
    if not depth.is_cuda:
        raise L.DynamoB200Error("backproject needs CUDA tensors (no CPU fallback)")
    return _BackprojectFn.apply(_prep(depth), _prep(inv_K))


class _ProjectFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, K, T, H, W):
        lib = L.load()
        B = points.shape[0]
        pix = torch.empty(B, H, W, 2, device=points.device)
        ego = torch.empty(B, 3, H * W, device=points.device)
        L.check(lib.dd_project_fwd(L.ptr(points), L.ptr(K), L.ptr(T), B, H, W, pix.data_ptr(), ego.data_ptr(), _stream()),
                "dd_project_fwd")
        ctx.save_for_backward(points, K, T)
        ctx.dims = (B, H, W)
        return pix, ego

    @staticmethod
    def backward(ctx, gpix, gego):
        lib = L.load()
        points, K, T = ctx.saved_tensors
        B, H, W = ctx.dims
        gp = torch.empty_like(points)
        gT = torch.empty(B, 4, 4, device=points.device) if (T is not None and ctx.needs_input_grad[2]) else None
        p = lambda t: t.contiguous().data_ptr() if t is not None else None
        L.check(lib.dd_project_bwd(points.data_ptr(), K.data_ptr(), p(T), p(gpix), p(gego), B, H, W, gp.data_ptr(), p(gT),
                                   _stream()), "dd_project_bwd")
        return gp, None, gT, None, None


def project3d(points, K, T, H, W):
    if not points.is_cuda:
        raise L.DynamoB200Error("project3d needs CUDA tensors (no CPU fallback)")
    return _ProjectFn.apply(_prep(points), _prep(K), _prep(T), int(H), int(W))


class _SsimFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        lib = L.load()
        H, W = x.shape[-2:]
        BC = x.numel() // (H * W)
        out = torch.empty_like(x)
        L.check(lib.dd_ssim_fwd(x.data_ptr(), y.data_ptr(), BC, H, W, out.data_ptr(), _stream()), "dd_ssim_fwd")
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    def backward(ctx, go):
        lib = L.load()
        x, y = ctx.saved_tensors
        H, W = x.shape[-2:]
        BC = x.numel() // (H * W)
        go = go.contiguous()
        gx = gy = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            L.check(lib.dd_ssim_bwd(x.data_ptr(), y.data_ptr(), go.data_ptr(), BC, H, W, gx.data_ptr(), _stream()), "dd_ssim_bwd")
        if ctx.needs_input_grad[1]:   # SSIM(x, y) is symmetric in its arguments
            gy = torch.empty_like(y)
            L.check(lib.dd_ssim_bwd(y.data_ptr(), x.data_ptr(), go.data_ptr(), BC, H, W, gy.data_ptr(), _stream()), "dd_ssim_bwd")
        return gx, gy


def ssim(x, y):
    if not x.is_cuda:
        raise L.DynamoB200Error("ssim needs CUDA tensors (no CPU fallback)")
    return _SsimFn.apply(_prep(x), _prep(y))


def ground_score(points, w, row0, tol):
    """counts (K,) int32 of ground points within `tol` of each plane hypothesis (tools.py:113-139); no gradient."""
    if not points.is_cuda:
        raise L.DynamoB200Error("ground_score needs CUDA tensors (no CPU fallback)")
    lib = L.load()
    points, w = _prep(points.detach()), _prep(w.detach().reshape(-1, 3))
    B, _, H, W = points.shape
    K = w.shape[0]
    counts = torch.empty(K, dtype=torch.int32, device=points.device)
    L.check(lib.dd_ground_score(points.data_ptr(), w.data_ptr(), B, H, W, int(row0), K, float(tol), counts.data_ptr(), _stream()),
            "dd_ground_score")
    return counts


# ---------------------------------------------------------------------------------------------
# pose head epilogue (pose_decoder.py:39-44, networks/layers.py:7-82)
# ---------------------------------------------------------------------------------------------


class _PoseMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        lib = L.load()
        B, Cc, h, w = x.shape
        out = torch.empty(B, Cc, device=x.device)
        L.check(lib.dd_pose_mean_fwd(x.data_ptr(), B * Cc, h * w, float(scale), out.data_ptr(), _stream()), "dd_pose_mean_fwd")
        ctx.dims = (B, Cc, h, w, float(scale))
        return out

    @staticmethod
    def backward(ctx, go):
        lib = L.load()
        B, Cc, h, w, scale = ctx.dims
        gx = torch.empty(B, Cc, h, w, device=go.device)
        L.check(lib.dd_pose_mean_bwd(go.contiguous().data_ptr(), B * Cc, h * w, scale, gx.data_ptr(), _stream()), "dd_pose_mean_bwd")
        return gx, None


def spatial_mean_scaled(x, scale):
    """scale * x.mean(3).mean(2) for (B,C,h,w) CUDA tensors."""
    if not x.is_cuda:
        raise L.DynamoB200Error("spatial_mean_scaled needs CUDA tensors (no CPU fallback)")
    return _PoseMeanFn.apply(_prep(x), scale)


class _PoseMatrixFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, aa, tr, invert):
        lib = L.load()
        B = aa.shape[0]
        T = torch.empty(B, 4, 4, device=aa.device)
        L.check(lib.dd_pose_matrix_fwd(aa.data_ptr(), tr.data_ptr(), B, int(invert), T.data_ptr(), _stream()), "dd_pose_matrix_fwd")
        ctx.save_for_backward(aa, tr)
        ctx.invert = int(invert)
        return T

    @staticmethod
    def backward(ctx, gT):
        lib = L.load()
        aa, tr = ctx.saved_tensors
        B = aa.shape[0]
        gaa, gtr = torch.empty_like(aa), torch.empty_like(tr)
        L.check(lib.dd_pose_matrix_bwd(aa.data_ptr(), tr.data_ptr(), gT.contiguous().data_ptr(), B, ctx.invert, gaa.data_ptr(),
                                       gtr.data_ptr(), _stream()), "dd_pose_matrix_bwd")
        return gaa, gtr, None


def pose_matrix(axisangle, translation, invert):
    """(B,1,3) axis-angle / translation -> (B,4,4) (transformation_from_parameters)."""
    B = axisangle.shape[0]
    return _PoseMatrixFn.apply(_prep(axisangle.reshape(B, 3)), _prep(translation.reshape(B, 3)), bool(invert))


class _LinearFn(torch.autograd.Function):
    """F.linear on the tcgen05 tensor cores at fp32 accuracy (dd_linear_fwd / dd_linear_bwd, csrc/linear_tc.cu)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = _prep(x).reshape(-1, x.shape[-1])
        w = _prep(weight)
        b = _prep(bias)
        M, K = x2.shape
        N = w.shape[0]
        y = torch.empty((M, N), device=x.device, dtype=torch.float32)
        with _timed("linear_fwd", meta=(2.0 * M * K * N, 4.0 * (M * K + N * K + M * N)), detail=True):
            lib = L.load()
            ws = _workspace(lib.dd_linear_workspace_bytes(M, K, N), x.device)
            L.check(lib.dd_linear_fwd(L.ptr(x2), L.ptr(w), L.ptr(b), M, K, N, L.ptr(y), L.ptr(ws), ws.numel(), _stream()), "dd_linear_fwd")
        ctx.save_for_backward(x2, w)
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        return y.reshape(x.shape[:-1] + (N,))

    @staticmethod
    def backward(ctx, g):
        x2, w = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        g2 = _prep(g).reshape(M, N)
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        gx = torch.empty((M, K), device=g.device, dtype=torch.float32) if need_x else None
        gw = torch.empty((N, K), device=g.device, dtype=torch.float32) if (need_w or need_b) else None
        gb = torch.empty((N,), device=g.device, dtype=torch.float32) if need_b else None
        flops = 2.0 * M * K * N * (int(bool(need_x)) + int(gw is not None))
        nbytes = 4.0 * ((M * N + N * K + M * K) * int(bool(need_x)) + (M * N + M * K + N * K) * int(gw is not None))
        with _timed("linear_bwd", meta=(flops, nbytes), detail=True):
            lib = L.load()
            ws = _workspace(lib.dd_linear_workspace_bytes(M, K, N), g.device)
            L.check(lib.dd_linear_bwd(L.ptr(x2), L.ptr(w), L.ptr(g2), M, K, N, L.ptr(gx), L.ptr(gw), L.ptr(gb), L.ptr(ws), ws.numel(),
                                      _stream()), "dd_linear_bwd")
        return (gx.reshape(ctx.x_shape) if need_x else None), (gw if need_w else None), gb


def linear(x, weight, bias=None):
    """y = x @ weight.T + bias over the last dimension of x (any leading shape); CUDA fp32 only."""
    return _LinearFn.apply(x, weight, bias)


class _ToNHWCFn(torch.autograd.Function):
    """(B,C,H,W) -> contiguous (B,H,W,C) (dd_nchw_to_nhwc); the gradient is the inverse transpose."""

    @staticmethod
    def forward(ctx, x):
        x = _prep(x)
        B, C, H, W = x.shape
        out = torch.empty((B, H, W, C), device=x.device, dtype=torch.float32)
        L.check(L.load().dd_nchw_to_nhwc(L.ptr(x), B, C, H * W, L.ptr(out), _stream()), "dd_nchw_to_nhwc")
        return out

    @staticmethod
    def backward(ctx, g):
        g = _prep(g)
        B, H, W, C = g.shape
        gx = torch.empty((B, C, H, W), device=g.device, dtype=torch.float32)
        L.check(L.load().dd_nhwc_to_nchw(L.ptr(g), B, C, H * W, L.ptr(gx), _stream()), "dd_nhwc_to_nchw")
        return gx


def nchw_to_nhwc(x):
    """x.permute(0, 2, 3, 1).contiguous() as one tiled transpose (and one for its gradient)."""
    return _ToNHWCFn.apply(x)


class _BlockTailFn(torch.autograd.Function):
    """out (B,C,H,W) = x + scale[b] * gamma[c] * y (B,H,W,C): layer scale + stochastic depth + residual of a Lite-Mono block."""

    @staticmethod
    def forward(ctx, x, y, gamma, scale):
        x, y, gamma, scale = _prep(x), _prep(y), _prep(gamma), _prep(scale)
        B, C, H, W = x.shape
        if tuple(y.shape) != (B, H, W, C):
            raise L.DynamoB200Error(f"block_tail: y {tuple(y.shape)} does not match x {tuple(x.shape)}")
        out = torch.empty_like(x)
        L.check(L.load().dd_block_tail_fwd(L.ptr(x), L.ptr(y), L.ptr(gamma), L.ptr(scale), B, C, H * W, L.ptr(out), _stream()),
                "dd_block_tail_fwd")
        ctx.save_for_backward(y, gamma, scale)
        ctx.dims = (B, C, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        y, gamma, scale = ctx.saved_tensors
        B, C, H, W = ctx.dims
        g = _prep(g)
        need_y = ctx.needs_input_grad[1]
        need_gamma = gamma is not None and ctx.needs_input_grad[2]
        gy = torch.empty_like(y) if need_y else None
        gg = torch.empty_like(gamma) if need_gamma else None
        if need_y or need_gamma:
            L.check(L.load().dd_block_tail_bwd(L.ptr(g), L.ptr(y), L.ptr(gamma), L.ptr(scale), B, C, H * W, L.ptr(gy), L.ptr(gg),
                                               _stream()), "dd_block_tail_bwd")
        return (g if ctx.needs_input_grad[0] else None), gy, gg, None


def block_tail(x, y, gamma=None, scale=None):
    """x + (scale[:, None, None, None] * gamma * y).permute(0, 3, 1, 2) with y in (B,H,W,C); scale is not differentiated."""
    return _BlockTailFn.apply(x, y, gamma, scale)


# ---------------------------------------------------------------------------------------------
# training-mode BatchNorm2d (+ exact GELU) of the Lite-Mono encoder (networks/depth_encoder.py:113-122, :194/:208)
# ---------------------------------------------------------------------------------------------
class _BatchNormGeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, gelu):
        x, weight, bias = _prep(x), _prep(weight), _prep(bias)
        B, C, H, W = x.shape
        lib = L.load()
        y = torch.empty_like(x)
        stats = torch.empty((2, C), device=x.device, dtype=torch.float32)
        ws = _workspace(lib.dd_bn_workspace_bytes(C), x.device)
        L.check(lib.dd_bn_gelu_fwd(L.ptr(x), B, C, H * W, L.ptr(weight), L.ptr(bias), float(eps), float(momentum), int(gelu), L.ptr(y),
                                   L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(running_mean), L.ptr(running_var), L.ptr(ws), ws.numel(),
                                   _stream()), "dd_bn_gelu_fwd")
        ctx.save_for_backward(x, weight, bias, stats)
        ctx.gelu = int(gelu)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, bias, stats = ctx.saved_tensors
        B, C, H, W = x.shape
        g = _prep(g)
        lib = L.load()
        need_x = ctx.needs_input_grad[0]
        need_w = weight is not None and ctx.needs_input_grad[1]
        need_b = bias is not None and ctx.needs_input_grad[2]
        if not (need_x or need_w or need_b):
            return (None,) * 8
        gx = torch.empty_like(x) if need_x else None
        gw = torch.empty_like(weight) if need_w else None
        gb = torch.empty_like(bias) if need_b else None
        ws = _workspace(lib.dd_bn_workspace_bytes(C), x.device)
        L.check(lib.dd_bn_gelu_bwd(L.ptr(x), L.ptr(g), B, C, H * W, L.ptr(weight), L.ptr(bias), L.ptr(stats[0]), L.ptr(stats[1]),
                                   ctx.gelu, L.ptr(gx), L.ptr(gw), L.ptr(gb), L.ptr(ws), ws.numel(), _stream()), "dd_bn_gelu_bwd")
        return gx, gw, gb, None, None, None, None, None


def batch_norm_gelu(x, bn, gelu=False):
    """`gelu(bn(x))` (or `bn(x)`) for an nn.BatchNorm2d in training mode on an NCHW CUDA tensor as two streaming passes
    (csrc/batchnorm.cu); running statistics and num_batches_tracked are updated exactly as nn.BatchNorm2d.forward does.
    Eval mode (running statistics) and modules without batch statistics stay on torch."""
    if not bn.training or not bn.track_running_stats or x.dim() != 4:
        y = bn(x)
        return torch.nn.functional.gelu(y) if gelu else y
    momentum = bn.momentum
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if momentum is None:   # cumulative moving average
            momentum = 1.0 / float(bn.num_batches_tracked)
    return _BatchNormGeluFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, 0.0 if momentum is None else momentum,
                                  bn.eps, gelu)


# ---------------------------------------------------------------------------------------------
# channels_last LayerNorm of the LGFI blocks (networks/depth_encoder.py:90-104)
# ---------------------------------------------------------------------------------------------
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x, weight, bias = _prep(x), _prep(weight), _prep(bias)
        C_ = x.shape[-1]
        M = x.numel() // C_
        y = torch.empty_like(x)
        stats = torch.empty((2, M), device=x.device, dtype=torch.float32)
        L.check(L.load().dd_layernorm_fwd(L.ptr(x), M, C_, L.ptr(weight), L.ptr(bias), float(eps), L.ptr(y), L.ptr(stats[0]),
                                          L.ptr(stats[1]), _stream()), "dd_layernorm_fwd")
        ctx.save_for_backward(x, weight, stats)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, stats = ctx.saved_tensors
        C_ = x.shape[-1]
        M = x.numel() // C_
        g = _prep(g)
        lib = L.load()
        need_x = ctx.needs_input_grad[0]
        need_w = weight is not None and ctx.needs_input_grad[1]
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        if not (need_x or need_w or need_b):
            return None, None, None, None
        gx = torch.empty_like(x) if need_x else None
        gw = torch.empty(C_, device=x.device, dtype=torch.float32) if need_w else None
        gb = torch.empty(C_, device=x.device, dtype=torch.float32) if need_b else None
        ws = _workspace(lib.dd_layernorm_workspace_bytes(C_), x.device) if (need_w or need_b) else None
        L.check(lib.dd_layernorm_bwd(L.ptr(x), L.ptr(g), M, C_, L.ptr(weight), L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(gx), L.ptr(gw),
                                     L.ptr(gb), L.ptr(ws), ws.numel() if ws is not None else 0, _stream()), "dd_layernorm_bwd")
        return gx, gw, gb, None


def layer_norm(x, weight=None, bias=None, eps=1e-6):
    """F.layer_norm(x, (C,), weight, bias, eps) over the last dimension of a CUDA tensor, C a multiple of 4 and <= 512."""
    return _LayerNormFn.apply(x, weight, bias, eps)


# ---------------------------------------------------------------------------------------------
# depth-wise dilated 3x3 convolution of the DilatedConv blocks (networks/depth_encoder.py:148-168)
# ---------------------------------------------------------------------------------------------
DWCONV_DILATIONS = (1, 2, 3, 4, 6)


class _DwConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, dilation):
        x, weight = _prep(x), _prep(weight)
        B, C_, H, W = x.shape
        if tuple(weight.shape) != (C_, 1, 3, 3):
            raise L.DynamoB200Error(f"dwconv3x3: weight {tuple(weight.shape)} does not match {C_} channels")
        y = torch.empty_like(x)
        L.check(L.load().dd_dwconv3x3_fwd(L.ptr(x), L.ptr(weight), B, C_, H, W, int(dilation), 0, L.ptr(y), _stream()), "dd_dwconv3x3_fwd")
        ctx.save_for_backward(x, weight)
        ctx.dilation = int(dilation)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        B, C_, H, W = x.shape
        g = _prep(g)
        lib = L.load()
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            L.check(lib.dd_dwconv3x3_fwd(L.ptr(g), L.ptr(weight), B, C_, H, W, ctx.dilation, 1, L.ptr(gx), _stream()), "dd_dwconv3x3_fwd")
        if ctx.needs_input_grad[1]:
            gw = torch.empty_like(weight)
            ws = _workspace(lib.dd_dwconv3x3_workspace_bytes(C_), x.device)
            L.check(lib.dd_dwconv3x3_wgrad(L.ptr(x), L.ptr(g), B, C_, H, W, ctx.dilation, L.ptr(gw), L.ptr(ws), ws.numel(), _stream()),
                    "dd_dwconv3x3_wgrad")
        return gx, gw, None


def dwconv3x3(x, weight, dilation=1):
    """F.conv2d(x, weight, None, 1, dilation, dilation, groups=C) for a (C,1,3,3) weight on an NCHW CUDA tensor (W % 4 == 0)."""
    return _DwConvFn.apply(x, weight, dilation)


# ---------------------------------------------------------------------------------------------
# nn.MaxPool2d(3, 2, 1) of the ResNet trunks on channels_last activations (networks/resnet_encoder.py:18,:130)
# ---------------------------------------------------------------------------------------------
def _nhwc_ptr(t):
    """Device pointer of a 4-D fp32 CUDA tensor that is dense in channels_last order."""
    if not t.is_cuda:
        raise L.DynamoB200Error("dynamo_b200 ops need CUDA tensors (no CPU fallback)")
    if t.dtype != torch.float32 or t.dim() != 4 or not t.is_contiguous(memory_format=torch.channels_last):
        raise L.DynamoB200Error("expected a channels_last fp32 (B,C,H,W) tensor")
    return t.data_ptr()


class _MaxPoolNHWCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if x.is_cuda and x.dim() == 4:
            x = x.float().contiguous(memory_format=torch.channels_last)
        B, C_, H, W = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((B, C_, Ho, Wo), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        arg = torch.empty((B, Ho, Wo, C_), device=x.device, dtype=torch.uint8)
        L.check(L.load().dd_maxpool3x3s2_nhwc_fwd(_nhwc_ptr(x), B, H, W, C_, _nhwc_ptr(y), L.ptr(arg), _stream()), "dd_maxpool3x3s2_nhwc_fwd")
        ctx.save_for_backward(arg)
        ctx.dims = (B, C_, H, W)
        return y

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        B, C_, H, W = ctx.dims
        g = g.float().contiguous(memory_format=torch.channels_last)
        gx = torch.empty((B, C_, H, W), device=g.device, dtype=torch.float32, memory_format=torch.channels_last)
        L.check(L.load().dd_maxpool3x3s2_nhwc_bwd(_nhwc_ptr(g), L.ptr(arg), B, H, W, C_, _nhwc_ptr(gx), _stream()), "dd_maxpool3x3s2_nhwc_bwd")
        return gx


def maxpool3x3s2(x):
    """F.max_pool2d(x, 3, 2, 1) on a channels_last CUDA tensor (C % 4 == 0); the result is channels_last as well."""
    return _MaxPoolNHWCFn.apply(x)


# ---------------------------------------------------------------------------------------------
# channels_last (ResNet trunks) -> NCHW (decoders) hand-over
# ---------------------------------------------------------------------------------------------
class _ToNCHWFn(torch.autograd.Function):
    """A channels_last (B,C,H,W) activation as an NCHW-contiguous tensor, and its gradient back in channels_last, each as one
    tiled transpose (dd_nhwc_to_nchw / dd_nchw_to_nhwc) instead of ATen's generic strided copy -- once per feature map, not
    once per consumer."""

    @staticmethod
    def forward(ctx, x):
        B, C_, H, W = x.shape
        out = torch.empty((B, C_, H, W), device=x.device, dtype=torch.float32)
        L.check(L.load().dd_nhwc_to_nchw(_nhwc_ptr(x), B, C_, H * W, L.ptr(out), _stream()), "dd_nhwc_to_nchw")
        return out

    @staticmethod
    def backward(ctx, g):
        g = _prep(g)
        B, C_, H, W = g.shape
        gx = torch.empty((B, C_, H, W), device=g.device, dtype=torch.float32, memory_format=torch.channels_last)
        L.check(L.load().dd_nchw_to_nhwc(L.ptr(g), B, C_, H * W, _nhwc_ptr(gx), _stream()), "dd_nchw_to_nhwc")
        return gx


def to_nchw(x):
    """NCHW-contiguous view/copy of a 4-D CUDA fp32 tensor; tensors that already are NCHW-contiguous pass through."""
    if x.dim() != 4 or x.is_contiguous() or not x.is_cuda or x.dtype != torch.float32:
        return x
    if not x.is_contiguous(memory_format=torch.channels_last):
        return x.contiguous()
    return _ToNCHWFn.apply(x)


# ---------------------------------------------------------------------------------------------
# cross-covariance attention core of the LGFI blocks (networks/depth_encoder.py:63-83, between qkv and proj)
# ---------------------------------------------------------------------------------------------
XCA_HEAD_DIMS = (8, 16, 28)


class _XcaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, temperature, heads):
        qkv, temp = _prep(qkv), _prep(temperature).reshape(-1)
        B, N, C3 = qkv.shape
        C_ = C3 // 3
        D = C_ // heads
        lib = L.load()
        out = torch.empty((B, N, C_), device=qkv.device, dtype=torch.float32)
        stats = torch.empty((2 * B * C_ * D + 2 * B * C_,), device=qkv.device, dtype=torch.float32)   # attn | scores | rq | rk
        ws = _workspace(lib.dd_xca_workspace_bytes(B, N, C_, heads), qkv.device)
        n1 = B * C_ * D
        attn, scores, rq, rk = stats[:n1], stats[n1:2 * n1], stats[2 * n1:2 * n1 + B * C_], stats[2 * n1 + B * C_:]
        L.check(lib.dd_xca_fwd(L.ptr(qkv), L.ptr(temp), B, N, C_, heads, L.ptr(out), L.ptr(attn), L.ptr(scores), L.ptr(rq), L.ptr(rk),
                               L.ptr(ws), ws.numel(), _stream()), "dd_xca_fwd")
        ctx.save_for_backward(qkv, temp, stats)
        ctx.dims = (B, N, C_, heads, tuple(temperature.shape))
        return out

    @staticmethod
    def backward(ctx, g):
        qkv, temp, stats = ctx.saved_tensors
        B, N, C_, heads, tshape = ctx.dims
        D = C_ // heads
        g = _prep(g)
        lib = L.load()
        n1 = B * C_ * D
        attn, scores, rq, rk = stats[:n1], stats[n1:2 * n1], stats[2 * n1:2 * n1 + B * C_], stats[2 * n1 + B * C_:]
        gqkv = torch.empty_like(qkv)
        gt_part = torch.empty((B, heads, D), device=g.device, dtype=torch.float32)
        ws = _workspace(lib.dd_xca_workspace_bytes(B, N, C_, heads), g.device)
        L.check(lib.dd_xca_bwd(L.ptr(qkv), L.ptr(temp), L.ptr(g), L.ptr(attn), L.ptr(scores), L.ptr(rq), L.ptr(rk), B, N, C_, heads,
                               L.ptr(gqkv), L.ptr(gt_part), L.ptr(ws), ws.numel(), _stream()), "dd_xca_bwd")
        gtemp = gt_part.sum((0, 2)).reshape(tshape) if ctx.needs_input_grad[1] else None
        return gqkv, gtemp, None


def xca_core(qkv, temperature, heads):
    """(softmax((q^ @ k^T) * temperature) @ v) of XCA.forward for the (B,N,3C) output of its qkv layer -> (B,N,C)."""
    return _XcaFn.apply(qkv, temperature, heads)


# ---------------------------------------------------------------------------------------------
# channels_last BatchNorm2d + residual + activation (ResNet BasicBlock pattern, Lite-Mono stem)
# ---------------------------------------------------------------------------------------------
BN_ACTS = {"none": 0, "relu": 1, "gelu": 2}


def bn_nhwc_supported(C_):
    return C_ % 4 == 0 and C_ // 4 <= 256 and 256 % (C_ // 4) == 0


class _BnActNHWCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, momentum, eps, act):
        x = x.float().contiguous(memory_format=torch.channels_last) if x.is_cuda else x
        if residual is not None:
            residual = residual.float().contiguous(memory_format=torch.channels_last)
        weight, bias = _prep(weight), _prep(bias)
        B, C_, H, W = x.shape
        lib = L.load()
        y = torch.empty((B, C_, H, W), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        stats = torch.empty((2, C_), device=x.device, dtype=torch.float32)
        ws = _workspace(lib.dd_bn_nhwc_workspace_bytes(C_), x.device)
        L.check(lib.dd_bn_act_nhwc_fwd(_nhwc_ptr(x), _nhwc_ptr(residual) if residual is not None else None, B * H * W, C_, L.ptr(weight),
                                       L.ptr(bias), float(eps), float(momentum), int(act), _nhwc_ptr(y), L.ptr(stats[0]), L.ptr(stats[1]),
                                       L.ptr(running_mean), L.ptr(running_var), L.ptr(ws), ws.numel(), _stream()), "dd_bn_act_nhwc_fwd")
        ctx.save_for_backward(x, y if act == 1 else None, weight, bias, stats)
        ctx.act, ctx.has_res = int(act), residual is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, y, weight, bias, stats = ctx.saved_tensors
        B, C_, H, W = x.shape
        g = g.float().contiguous(memory_format=torch.channels_last)
        lib = L.load()
        need_x, need_r = ctx.needs_input_grad[0], ctx.has_res and ctx.needs_input_grad[1]
        need_w = weight is not None and ctx.needs_input_grad[2]
        need_b = bias is not None and ctx.needs_input_grad[3]
        if not (need_x or need_r or need_w or need_b):
            return (None,) * 9
        gx = torch.empty_like(x) if need_x else None                      # (empty_like keeps channels_last)
        gr = torch.empty_like(x) if need_r else None
        gw = torch.empty_like(weight) if need_w else None
        gb = torch.empty_like(bias) if need_b else None
        ws = _workspace(lib.dd_bn_nhwc_workspace_bytes(C_), x.device)
        L.check(lib.dd_bn_act_nhwc_bwd(_nhwc_ptr(x), _nhwc_ptr(y) if y is not None else None, _nhwc_ptr(g), B * H * W, C_, L.ptr(weight),
                                       L.ptr(bias), L.ptr(stats[0]), L.ptr(stats[1]), ctx.act, _nhwc_ptr(gx) if gx is not None else None,
                                       _nhwc_ptr(gr) if gr is not None else None, L.ptr(gw), L.ptr(gb), L.ptr(ws), ws.numel(), _stream()),
                "dd_bn_act_nhwc_bwd")
        return gx, gr, gw, gb, None, None, None, None, None


def bn_act_nhwc(x, bn, act="none", residual=None):
    """act(bn(x) + residual) for an nn.BatchNorm2d in training mode on a channels_last CUDA tensor (csrc/batchnorm_nhwc.cu);
    running statistics / num_batches_tracked updated as nn.BatchNorm2d.forward does.  Eval mode and unsupported channel counts
    take the torch formulation."""
    if not bn.training or not bn.track_running_stats or x.dim() != 4 or not x.is_cuda or not bn_nhwc_supported(x.shape[1]):
        y = bn(x)
        if residual is not None:
            y = y + residual
        return torch.relu(y) if act == "relu" else (torch.nn.functional.gelu(y) if act == "gelu" else y)
    momentum = bn.momentum
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked)
    return _BnActNHWCFn.apply(x, residual, bn.weight, bn.bias, bn.running_mean, bn.running_var, 0.0 if momentum is None else momentum,
                              bn.eps, BN_ACTS[act])
