"""Depth decoders (reference: networks/depth_decoder.py).  Same constructor arguments, module tree and
state_dict keys; each up-conv block (x2 up-sampling -> skip concat -> reflect pad -> 3x3 conv -> ELU)
is ONE fused kernel launch instead of five ATen/cuDNN ops."""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from dd_b200.functional import resize_bilinear
from .layers import ConvBlock, Conv3x3


class DepthDecoder(nn.Module):
    """Monodepth2 decoder: nearest x2 up-sampling, disparities at `scales` (depth_decoder.py:10-55)."""

    def __init__(self, num_ch_enc, scales=range(4), num_output_channels=1, use_skips=True):
        super().__init__()
        self.num_output_channels = num_output_channels
        self.use_skips = use_skips
        self.upsample_mode = "nearest"
        self.scales = scales
        self.num_ch_enc = num_ch_enc
        self.num_ch_dec = np.array([16, 32, 64, 128, 256])
        for i in range(4, -1, -1):
            cin = self.num_ch_enc[-1] if i == 4 else self.num_ch_dec[i + 1]
            setattr(self, f"upconv_{i}_0", ConvBlock(cin, self.num_ch_dec[i]))
            cin = self.num_ch_dec[i] + (self.num_ch_enc[i - 1] if (self.use_skips and i > 0) else 0)
            setattr(self, f"upconv_{i}_1", ConvBlock(cin, self.num_ch_dec[i]))
        for s in self.scales:
            setattr(self, f"dispconv_{s}", Conv3x3(self.num_ch_dec[s], self.num_output_channels))
        self.sigmoid = nn.Sigmoid()

    def forward(self, input_features):
        self.outputs = {}
        x = input_features[-1]
        for i in range(4, -1, -1):
            x = getattr(self, f"upconv_{i}_0")(x)
            skip = input_features[i - 1] if (self.use_skips and i > 0) else None
            x = getattr(self, f"upconv_{i}_1")(x, skip=skip, up="nearest")
            if i in self.scales:
                self.outputs[("disp", i)] = getattr(self, f"dispconv_{i}")(x, act="sigmoid")
        return self.outputs


class LiteDepthDecoder(nn.Module):
    """Lite-Mono decoder: bilinear x2 up-sampling, disparity = sigmoid(bilinear x2 (dispconv))
    (depth_decoder.py:58-115); parameters are registered through `self.decoder` in the reference's order."""

    def __init__(self, num_ch_enc, scales=range(4), num_output_channels=1, use_skips=True):
        super().__init__()
        self.num_output_channels = num_output_channels
        self.use_skips = use_skips
        self.upsample_mode = "bilinear"
        self.scales = scales
        self.num_ch_enc = num_ch_enc
        self.num_ch_dec = (self.num_ch_enc / 2).astype("int")
        self.convs = OrderedDict()
        for i in range(2, -1, -1):
            cin = self.num_ch_enc[-1] if i == 2 else self.num_ch_dec[i + 1]
            self.convs[("upconv", i, 0)] = ConvBlock(cin, self.num_ch_dec[i])
            cin = self.num_ch_dec[i] + (self.num_ch_enc[i - 1] if (self.use_skips and i > 0) else 0)
            self.convs[("upconv", i, 1)] = ConvBlock(cin, self.num_ch_dec[i])
        for s in self.scales:
            self.convs[("dispconv", s)] = Conv3x3(self.num_ch_dec[s], self.num_output_channels)
        self.decoder = nn.ModuleList(list(self.convs.values()))
        self.sigmoid = nn.Sigmoid()
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.trunc_normal_(m.weight, std=0.02)   # == timm trunc_normal_
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def forward(self, input_features):
        self.outputs = {}
        x = input_features[-1]
        for i in range(2, -1, -1):
            x = self.convs[("upconv", i, 0)](x)
            skip = input_features[i - 1] if (self.use_skips and i > 0) else None
            x = self.convs[("upconv", i, 1)](x, skip=skip, up="bilinear")
            if i in self.scales:
                raw = self.convs[("dispconv", i)](x)
                self.outputs[("disp", i)] = resize_bilinear(raw, (2 * raw.shape[-2], 2 * raw.shape[-1]), sigmoid=True)
        return self.outputs
