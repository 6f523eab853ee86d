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
    aux: Dict = field(default_factory=dict)  # filled by the forward pass: (name, frame_index, level) -> tensor

    @property
    def flags(self):
        return (L.DD_FLAG_CMPFLOW if self.cmpflow else 0) | (L.DD_FLAG_MOTMASK if self.motmask else 0) | \
               (L.DD_FLAG_AUTOMASK if self.automask else 0)


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
        L.check(lib.dd_warp_photo_fwd(C.byref(desc), C.byref(aux), sums.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
                "dd_warp_photo_fwd")

        ctx.cfg, ctx.nF = cfg, nF
        ctx.save_for_backward(*tensors)
        ctx.saved_resid = {k: v for k, v in out.items() if k[0] == "resid"}
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
        for (name, f, i), t in ctx.saved_resid.items():
            saved.resid[i][f] = t.data_ptr()
        gs = grad_sums.contiguous().float()
        ws_bytes = lib.dd_warp_photo_workspace_bytes(C.byref(desc))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
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
