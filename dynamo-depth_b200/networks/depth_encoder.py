"""Lite-Mono-8M depth encoder (reference: networks/depth_encoder.py:293-431).

Kept in PyTorch (encoder contractions are the tensor-core / cuDNN part of the step).  The module
tree reproduces the reference's attribute names so its 263 state tensors load by key:
  downsample_layers.0.{0,1,2}.conv / .bn_gelu.bn, stem2.0.conv, downsample_layers.{1,2}.0.conv,
  stages.i.j.{ddwconv.conv, bn1, norm, pwconv1, pwconv2, gamma} for the dilated blocks and
  stages.i.last.{pos_embd.token_projection, norm_xca, gamma_xca, xca.{temperature,qkv,proj}, norm,
  pwconv1, pwconv2, gamma} for the LGFI block closing each stage.
timm is not a dependency: DropPath is implemented here with timm-0.6.13 semantics.
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class DropPath(nn.Module):
    """Stochastic depth per sample: x * Bernoulli(keep) / keep in training, identity in eval."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def sample(self, x):
        """The per-sample factor Bernoulli(keep) / keep in the shape (B, 1, ..) (one RNG draw of B numbers), or None."""
        if self.drop_prob == 0.0 or not self.training:
            return None
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0:
            mask.div_(keep)
        return mask

    def forward(self, x):
        mask = self.sample(x)
        return x if mask is None else x * mask


class _LinearTF32(torch.autograd.Function):
    """F.linear with the three GEMMs (forward, input gradient, weight gradient) on the TF32 tensor cores; the global
    matmul precision flag is only raised around these calls, so nothing else (e.g. the RANSAC least squares) changes."""

    @staticmethod
    def _tf32(fn):
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            return fn()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return _LinearTF32._tf32(lambda: F.linear(x, weight, bias))

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g2, x2 = g.reshape(-1, g.shape[-1]), x.reshape(-1, x.shape[-1])
        gx = _LinearTF32._tf32(lambda: g2 @ weight).reshape(x.shape) if ctx.needs_input_grad[0] else None
        gw = _LinearTF32._tf32(lambda: g2.t() @ x2) if ctx.needs_input_grad[1] else None
        gb = g2.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return gx, gw, gb


class EncoderLinear(nn.Linear):
    """nn.Linear of the encoder (same parameters / state_dict keys; reference networks/depth_encoder.py:58-60,
    :197-199, :243-245).  On CUDA the three contractions (forward, input gradient, weight gradient + bias gradient)
    run in the hand-written tcgen05 kernel of csrc/linear_tc.cu: 3xTF32 operand split with fp32 accumulation in TMEM,
    i.e. fp32 accuracy on the tensor cores (dd_b200.functional.linear).  `tf32 = True` (opt-in, options
    --encoder_tf32_linear) uses cuBLAS single-pass TF32 instead, `mode = "torch"` (options --encoder_linear torch)
    torch's fp32 SIMT matmul.  CPU tensors (oracle-side tests) always take F.linear."""
    tf32 = False
    mode = "tc3x"

    def forward(self, x):
        if x.is_cuda:
            if EncoderLinear.tf32:
                return _LinearTF32.apply(x, self.weight, self.bias)
            if EncoderLinear.mode == "tc3x":
                from dd_b200 import functional as DF
                return DF.linear(x, self.weight, self.bias)
        return F.linear(x, self.weight, self.bias)


class LayerNorm(nn.Module):
    """channels_last (default) or channels_first layer norm with learnable affine."""

    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_last"):
        super().__init__()
        if data_format not in ("channels_last", "channels_first"):
            raise NotImplementedError
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps, self.data_format = eps, data_format
        self.normalized_shape = (normalized_shape,)

    def forward(self, x):
        if self.data_format == "channels_last":
            if x.is_cuda and EncoderLinear.mode != "torch" and x.shape[-1] % 4 == 0 and x.shape[-1] <= 512:
                from dd_b200 import functional as DF   # row-per-sub-warp kernel (csrc/layernorm.cu)
                return DF.layer_norm(x, self.weight, self.bias, self.eps)
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        mu = x.mean(1, keepdim=True)
        var = (x - mu).pow(2).mean(1, keepdim=True)
        return self.weight[:, None, None] * ((x - mu) / torch.sqrt(var + self.eps)) + self.bias[:, None, None]


class PositionalEncodingFourier(nn.Module):
    """DETR-style 2-D sine/cosine position code projected to `dim` channels by a 1x1 convolution."""

    def __init__(self, hidden_dim=32, dim=768, temperature=10000):
        super().__init__()
        self.token_projection = nn.Conv2d(hidden_dim * 2, dim, kernel_size=1)
        self.scale = 2 * math.pi
        self.temperature, self.hidden_dim, self.dim = temperature, hidden_dim, dim

    def forward(self, B, H, W):
        dev = self.token_projection.weight.device
        ys = torch.arange(1, H + 1, dtype=torch.float32, device=dev).view(1, H, 1).expand(B, H, W)
        xs = torch.arange(1, W + 1, dtype=torch.float32, device=dev).view(1, 1, W).expand(B, H, W)
        eps = 1e-6
        ys = ys / (ys[:, -1:, :] + eps) * self.scale
        xs = xs / (xs[:, :, -1:] + eps) * self.scale
        k = torch.arange(self.hidden_dim, dtype=torch.float32, device=dev)
        freq = self.temperature ** (2 * torch.div(k, 2, rounding_mode="trunc") / self.hidden_dim)
        px, py = xs[..., None] / freq, ys[..., None] / freq
        px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
        py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
        return self.token_projection(torch.cat((py, px), dim=3).permute(0, 3, 1, 2))


class XCA(nn.Module):
    """Cross-covariance attention: softmax over the (d_h x d_h) channel covariance of L2-normalised q, k."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.qkv = EncoderLinear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = EncoderLinear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, N, C = x.shape
        if x.is_cuda and EncoderLinear.mode != "torch" and not (self.training and self.attn_drop.p > 0) and C % 32 == 0 and C <= 256:
            from dd_b200 import functional as DF
            if C // self.num_heads in DF.XCA_HEAD_DIMS:   # streaming Gram / softmax / apply kernels (csrc/xca.cu)
                return self.proj_drop(self.proj(DF.xca_core(self.qkv(x), self.temperature, self.num_heads)))
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 4, 1)   # (3,B,h,d,N)
        q, k, v = F.normalize(qkv[0], dim=-1), F.normalize(qkv[1], dim=-1), qkv[2]
        attn = self.attn_drop(((q @ k.transpose(-2, -1)) * self.temperature).softmax(dim=-1))
        out = (attn @ v).permute(0, 3, 1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(out))


class BNGELU(nn.Module):
    def __init__(self, nIn):
        super().__init__()
        self.bn = nn.BatchNorm2d(nIn, eps=1e-5)
        self.act = nn.GELU()

    def forward(self, x):
        if _fused(x):   # training-mode statistics + affine + GELU as two streaming passes (csrc/batchnorm.cu)
            from dd_b200 import functional as DF
            return DF.batch_norm_gelu(x, self.bn, gelu=True)
        return self.act(self.bn(x))


class Conv(nn.Module):
    def __init__(self, nIn, nOut, kSize, stride, padding=0, dilation=(1, 1), groups=1, bn_act=False, bias=False):
        super().__init__()
        self.bn_act = bn_act
        self.conv = nn.Conv2d(nIn, nOut, kernel_size=kSize, stride=stride, padding=padding, dilation=dilation,
                              groups=groups, bias=bias)
        if bn_act:
            self.bn_gelu = BNGELU(nOut)

    def forward(self, x):
        x = self.conv(x)
        if self.bn_act and _fused(x) and x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last):
            from dd_b200 import functional as DF   # channels_last stem: BN + GELU in NHWC (csrc/batchnorm_nhwc.cu)
            return DF.bn_act_nhwc(x, self.bn_gelu.bn, "gelu")
        return self.bn_gelu(x) if self.bn_act else x


class CDilated(nn.Module):
    def __init__(self, nIn, nOut, kSize, stride=1, d=1, groups=1, bias=False):
        super().__init__()
        self.conv = nn.Conv2d(nIn, nOut, kSize, stride=stride, padding=int((kSize - 1) / 2) * d, bias=bias, dilation=d,
                              groups=groups)

    def forward(self, x):
        c = self.conv
        if (_fused(x) and c.groups == c.in_channels == c.out_channels and c.kernel_size == (3, 3) and c.stride == (1, 1)
                and c.bias is None and x.shape[-1] % 4 == 0 and c.dilation[0] == c.dilation[1] and c.padding == c.dilation):
            from dd_b200 import functional as DF
            if c.dilation[0] in DF.DWCONV_DILATIONS:   # streaming depth-wise kernel (csrc/dwconv.cu)
                return DF.dwconv3x3(x, c.weight, c.dilation[0])
        return self.conv(x)


def _mlp_branch(block, x):
    """shared inverted-bottleneck tail: Linear(dim->6dim) - GELU - Linear(6dim->dim) - layer scale."""
    x = block.pwconv2(block.act(block.pwconv1(x)))
    return block.gamma * x if block.gamma is not None else x


def _fused(x):
    """CUDA tensors take the hand-written layout-glue kernels (csrc/lite_glue.cu) unless EncoderLinear.mode == "torch"."""
    return x.is_cuda and EncoderLinear.mode != "torch"


def _block_tail(block, x, y):
    """x + drop_path(gamma * y).permute(0, 3, 1, 2) for y in (B,H,W,C) as one kernel; the stochastic-depth factor is drawn
    exactly like DropPath.forward draws it (same shape, same RNG stream)."""
    from dd_b200 import functional as DF
    scale = block.drop_path.sample(x) if isinstance(block.drop_path, DropPath) else None
    return DF.block_tail(x, y, block.gamma, None if scale is None else scale.reshape(-1))


class DilatedConv(nn.Module):
    """One block of the consecutive-dilated-convolution module: depth-wise dilated 3x3, BN, MLP, residual."""

    def __init__(self, dim, k, dilation=1, stride=1, drop_path=0.0, layer_scale_init_value=1e-6, expan_ratio=6):
        super().__init__()
        self.ddwconv = CDilated(dim, dim, kSize=k, stride=stride, groups=dim, d=dilation)
        self.bn1 = nn.BatchNorm2d(dim)
        self.norm = LayerNorm(dim, eps=1e-6)   # present in the state dict, unused in forward (as in the reference)
        self.pwconv1 = EncoderLinear(dim, expan_ratio * dim)
        self.act = nn.GELU()
        self.pwconv2 = EncoderLinear(expan_ratio * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x):
        if _fused(x):
            from dd_b200 import functional as DF
            y = DF.nchw_to_nhwc(DF.batch_norm_gelu(self.ddwconv(x), self.bn1))
            return _block_tail(self, x, self.pwconv2(self.act(self.pwconv1(y))))
        y = self.bn1(self.ddwconv(x)).permute(0, 2, 3, 1)
        y = _mlp_branch(self, y).permute(0, 3, 1, 2)
        return x + self.drop_path(y)


class LGFI(nn.Module):
    """Local-global feature interaction: (optional position code) + XCA residual, then the MLP branch."""

    def __init__(self, dim, drop_path=0.0, layer_scale_init_value=1e-6, expan_ratio=6, use_pos_emb=True, num_heads=6,
                 qkv_bias=True, attn_drop=0.0, drop=0.0):
        super().__init__()
        self.dim = dim
        self.pos_embd = PositionalEncodingFourier(dim=dim) if use_pos_emb else None
        self.norm_xca = LayerNorm(dim, eps=1e-6)
        self.gamma_xca = nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None
        self.xca = XCA(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = EncoderLinear(dim, expan_ratio * dim)
        self.act = nn.GELU()
        self.pwconv2 = EncoderLinear(expan_ratio * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x):
        B, C, H, W = x.shape
        fused = _fused(x)
        if fused:
            from dd_b200 import functional as DF
            t = DF.nchw_to_nhwc(x).reshape(B, H * W, C)
        else:
            t = x.reshape(B, C, H * W).permute(0, 2, 1)
        if self.pos_embd is not None:
            # the position code does not depend on the image: on the fused path it is evaluated for ONE image and broadcast over
            # the batch (the reference builds B identical copies, networks/depth_encoder.py:257-259)
            pb = 1 if fused else B
            t = t + self.pos_embd(pb, H, W).reshape(pb, -1, t.shape[1]).permute(0, 2, 1)
        t = t + self.gamma_xca * self.xca(self.norm_xca(t))
        if fused:
            y = self.norm(t.reshape(B, H, W, C))
            return _block_tail(self, x, self.pwconv2(self.act(self.pwconv1(y))))
        y = _mlp_branch(self, self.norm(t.reshape(B, H, W, C))).permute(0, 3, 1, 2)
        return x + self.drop_path(y)


class AvgPool(nn.Module):
    def __init__(self, ratio):
        super().__init__()
        self.pool = nn.ModuleList([nn.AvgPool2d(3, stride=2, padding=1) for _ in range(ratio)])

    def forward(self, x):
        for p in self.pool:
            x = p(x)
        return x


class LiteMono(nn.Module):
    stem_channels_last = True
    def __init__(self, in_chans=3, model="lite-mono-8m", global_block=[1, 1, 1], global_block_type=["LGFI", "LGFI", "LGFI"],
                 drop_path_rate=0.2, layer_scale_init_value=1e-6, expan_ratio=6, heads=[8, 8, 8],
                 use_pos_embd_xca=[True, False, False], pretrained=True, **kwargs):
        super().__init__()
        assert model == "lite-mono-8m", "Only using lite-mono-8m"
        self.num_ch_enc = np.array([64, 128, 224])
        self.depth = [4, 4, 10]
        self.dims = [64, 128, 224]
        self.dilation = [[1, 2, 3], [1, 2, 3], [1, 2, 3, 1, 2, 3, 2, 4, 6]]
        assert all(g in ("None", "LGFI") for g in global_block_type)
        d0 = self.dims[0]

        self.downsample_layers = nn.ModuleList()
        self.downsample_layers.append(nn.Sequential(
            Conv(in_chans, d0, kSize=3, stride=2, padding=1, bn_act=True),
            Conv(d0, d0, kSize=3, stride=1, padding=1, bn_act=True),
            Conv(d0, d0, kSize=3, stride=1, padding=1, bn_act=True)))
        self.stem2 = nn.Sequential(Conv(d0 + 3, d0, kSize=3, stride=2, padding=1, bn_act=False))
        self.input_downsample = nn.ModuleList([AvgPool(i) for i in range(1, 5)])
        for i in range(2):
            self.downsample_layers.append(nn.Sequential(
                Conv(self.dims[i] * 2 + 3, self.dims[i + 1], kSize=3, stride=2, padding=1, bn_act=False)))

        rates = [r.item() for r in torch.linspace(0, drop_path_rate, sum(self.depth))]
        self.stages = nn.ModuleList()
        first = 0
        for i in range(3):
            blocks = []
            for j in range(self.depth[i]):
                if j > self.depth[i] - global_block[i] - 1:
                    if global_block_type[i] != "LGFI":
                        raise NotImplementedError
                    blocks.append(LGFI(dim=self.dims[i], drop_path=rates[first + j], expan_ratio=expan_ratio,
                                       use_pos_emb=use_pos_embd_xca[i], num_heads=heads[i],
                                       layer_scale_init_value=layer_scale_init_value))
                else:
                    blocks.append(DilatedConv(dim=self.dims[i], k=3, dilation=self.dilation[i][j], drop_path=rates[first + j],
                                              layer_scale_init_value=layer_scale_init_value, expan_ratio=expan_ratio))
            self.stages.append(nn.Sequential(*blocks))
            first += self.depth[i]

        self.apply(self._init_weights)
        if pretrained:
            self.load_pretrained_model(model)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        elif isinstance(m, (LayerNorm, nn.LayerNorm)):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)

    def load_pretrained_model(self, model_name, dev="cpu"):
        path = f"./ckpt/{model_name}-pretrain.pth"
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not found (no network access here: place the Lite-Mono ImageNet checkpoint "
                                    "there or use --weights_init scratch)")
        own = self.state_dict()
        own.update({k: v for k, v in torch.load(path, map_location=dev)["model"].items() if k in own and not k.startswith("norm")})
        self.load_state_dict(own)

    def forward_features(self, x):
        x = (x - 0.45) / 0.225
        pyramid = [pool(x) for pool in self.input_downsample]
        if _fused(x) and LiteMono.stem_channels_last:
            # the three 3x3 stem convolutions and stem2 run channels_last: cuDNN's tensor-core kernels are NHWC, an NCHW stem pays
            # a layout conversion of the 250 MB half-resolution maps in front of and behind every convolution
            from dd_b200 import functional as DF
            h = self.downsample_layers[0](x.contiguous(memory_format=torch.channels_last))
            x = DF.to_nchw(self.stem2(torch.cat((h, pyramid[0].contiguous(memory_format=torch.channels_last)), dim=1)))
        else:
            x = self.stem2(torch.cat((self.downsample_layers[0](x), pyramid[0]), dim=1))
        features, carry = [], [x]
        for i in range(3):
            if i > 0:
                carry.append(pyramid[i])
                x = self.downsample_layers[i](torch.cat(carry, dim=1))
                carry = [x]
            x = self.stages[i](x)
            carry.append(x)
            features.append(x)
        return features

    def forward(self, x):
        return self.forward_features(x)
