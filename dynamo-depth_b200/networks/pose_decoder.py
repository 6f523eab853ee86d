"""Pose head (reference: networks/pose_decoder.py): 1x1 squeeze + ReLU, two 3x3 zero-padded
convolutions + ReLU, 1x1 to 6 values per predicted frame, spatial mean, x0.01."""
from collections import OrderedDict

import torch
import torch.nn as nn

from dd_b200.functional import conv2d_fused, spatial_mean_scaled


class PoseDecoder(nn.Module):
    def __init__(self, num_ch_enc, num_input_features, num_frames_to_predict_for=None, stride=1):
        super().__init__()
        if stride != 1:
            raise NotImplementedError("PoseDecoder: only stride=1 (the value Model uses) is implemented")
        self.num_ch_enc = num_ch_enc
        self.num_input_features = num_input_features
        self.num_frames_to_predict_for = num_frames_to_predict_for or (num_input_features - 1)
        self.squeeze = nn.Conv2d(self.num_ch_enc[-1], 256, 1)
        self.pose0 = nn.Conv2d(num_input_features * 256, 256, 3, stride, 1)
        self.pose1 = nn.Conv2d(256, 256, 3, stride, 1)
        self.pose2 = nn.Conv2d(256, 6 * self.num_frames_to_predict_for, 1)
        self.convs = OrderedDict()
        self.net = nn.ModuleList([self.squeeze, self.pose0, self.pose1, self.pose2])   # aliases net.{0..3}.* in the state_dict
        self.relu = nn.ReLU()

    def forward(self, input_features):
        sq = [conv2d_fused(f[-1], self.squeeze.weight, self.squeeze.bias, ksize=1, act="relu") for f in input_features]
        x = sq[0] if len(sq) == 1 else torch.cat(sq, 1)
        x = conv2d_fused(x, self.pose0.weight, self.pose0.bias, ksize=3, pad="zero", act="relu")
        x = conv2d_fused(x, self.pose1.weight, self.pose1.bias, ksize=3, pad="zero", act="relu")
        x = conv2d_fused(x, self.pose2.weight, self.pose2.bias, ksize=1)
        out = spatial_mean_scaled(x, 0.01).view(-1, self.num_frames_to_predict_for, 1, 6)
        return out[..., :3], out[..., 3:]
