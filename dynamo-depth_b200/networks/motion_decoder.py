"""Scene-flow / motion-mask decoder (reference: networks/motion_decoder.py).  Coarse-to-fine residual
refinement of a motion field seeded by the ego-motion; per level: bilinear up-sampling to the
encoder feature's size, concat, two zero-padded 3x3 convolutions without non-linearity, a 1x1
reduction over cat(x1, x2) and a residual add -- executed by dd_resize_bilinear_* and dd_conv_*
(the concatenations and the residual add are fused into the convolution launches)."""
import torch
import torch.nn as nn

from dd_b200.functional import conv2d_fused, resize_bilinear


class MotionDecoder(nn.Module):
    def __init__(self, num_inp_feat, scales=4, num_input_images=2, inp_disp=True, out_dim=4):
        super().__init__()
        self.org_in_ch = num_input_images * (3 + int(inp_disp))
        self.num_inp_feat = num_inp_feat[::-1].tolist() + [self.org_in_ch]
        self.out_dim = out_dim
        self.scales = scales
        assert max(self.scales) < len(self.num_inp_feat)
        self._residual_translation = nn.Conv2d(6, self.out_dim, kernel_size=1, stride=1, padding=0)
        for ii, c in enumerate(self.num_inp_feat):
            setattr(self, f"refine_motion_conv{ii}", nn.Sequential(nn.Conv2d(c + self.out_dim, c, 3, 1, 1), nn.Conv2d(c, c, 3, 1, 1)))
            setattr(self, f"refine_motion_redu{ii}", nn.Conv2d(c * 2, self.out_dim, 1, 1))

    def _refine_motion_field(self, feats, field):
        """feats: encoder pyramid coarse -> fine; returns the refined field of every level."""
        levels = []
        for ii in range(len(self.num_inp_feat)):
            feat = feats[-1 - ii]
            up = resize_bilinear(field, feat.shape[-2:])                              # motion_decoder.py:38
            conv, redu = getattr(self, f"refine_motion_conv{ii}"), getattr(self, f"refine_motion_redu{ii}")
            x1 = conv2d_fused(up, conv[0].weight, conv[0].bias, x1=feat, ksize=3, pad="zero")      # cat(up, feat)
            x2 = conv2d_fused(x1, conv[1].weight, conv[1].bias, ksize=3, pad="zero")
            field = conv2d_fused(x1, redu.weight, redu.bias, x1=x2, residual=up, ksize=1)          # redu(cat) + up
            levels.append(field)
        return levels

    def forward(self, pose_feat, ego_motion):
        """pose_feat: [input (B,3N,H,W), feat/2, /4, /8, /16, /32]; ego_motion (B,6,1,1)."""
        self.outputs = pose_feat
        rt = self._residual_translation
        seed = conv2d_fused(100 * ego_motion, rt.weight, rt.bias, ksize=1)
        levels = self._refine_motion_field(pose_feat, seed)
        outputs = {}
        for scale in self.scales:
            m_raw = 0.01 * levels[len(self.num_inp_feat) - 1 - scale]
            if self.out_dim == 1:
                outputs[("motion_prob", scale)] = m_raw
                outputs[("motion_mask", scale)] = torch.sigmoid(m_raw)
            elif self.out_dim == 3:
                outputs[("complete_flow", scale)] = m_raw
            else:
                raise Exception(f"out_dim={self.out_dim} not excepted.")
        return outputs
