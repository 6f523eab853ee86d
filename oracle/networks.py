"""ORACLE (test infrastructure, never on the product path) -- plain-PyTorch CPU restatement of the
reference's decoders, pose head, model wiring and training step.

Functional style: every decoder is a function of (features, state_dict) using only F.conv2d / F.pad /
F.interpolate / elementwise ops, with the reference's state_dict keys, so the same weights drive the
reference, this oracle and the CUDA product.  Pinned by tests/golden/nets_*.npz (recorded from the
unmodified reference by oracle/gen_golden_nets.py).  Encoders are injected (they are ordinary PyTorch
modules in the reference and in the product alike).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import view_synthesis as vs


def conv3x3(x, w, b, reflect=True):
    """networks/layers.py:100-115 (ReflectionPad2d(1) | ZeroPad2d(1), then Conv2d 3x3)."""
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect" if reflect else "constant"), w, b)


def conv_block(x, sd, prefix):
    """networks/layers.py:85-97: Conv3x3 + ELU; keys '<prefix>.conv.conv.{weight,bias}'."""
    return F.elu(conv3x3(x, sd[prefix + ".conv.conv.weight"], sd[prefix + ".conv.conv.bias"]))


def depth_decoder_md2(feats, sd, scales):
    """networks/depth_decoder.py:40-55: nearest x2, skip concat, sigmoid(dispconv)."""
    out = {}
    x = feats[-1]
    for i in range(4, -1, -1):
        x = conv_block(x, sd, f"upconv_{i}_0")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if i > 0:
            x = torch.cat([x, feats[i - 1]], 1)
        x = conv_block(x, sd, f"upconv_{i}_1")
        if i in scales:
            out[("disp", i)] = torch.sigmoid(conv3x3(x, sd[f"dispconv_{i}.conv.weight"], sd[f"dispconv_{i}.conv.bias"]))
    return out


def lite_decoder_keys(scales):
    """LiteDepthDecoder registers its blocks through nn.ModuleList 'decoder' in construction order
    (networks/depth_decoder.py:71-88): upconv (2,0),(2,1),(1,0),(1,1),(0,0),(0,1), then dispconv per scale."""
    order = [("upconv", i, j) for i in (2, 1, 0) for j in (0, 1)] + [("dispconv", s) for s in scales]
    return {k: f"decoder.{n}" for n, k in enumerate(order)}


def depth_decoder_lite(feats, sd, scales):
    """networks/depth_decoder.py:99-115: bilinear x2, skip concat, sigmoid(bilinear x2(dispconv))."""
    keys = lite_decoder_keys(scales)
    out = {}
    x = feats[-1]
    for i in range(2, -1, -1):
        x = conv_block(x, sd, keys[("upconv", i, 0)])
        x = F.interpolate(x, scale_factor=2, mode="bilinear")
        if i > 0:
            x = torch.cat([x, feats[i - 1]], 1)
        x = conv_block(x, sd, keys[("upconv", i, 1)])
        if i in scales:
            k = keys[("dispconv", i)]
            raw = conv3x3(x, sd[k + ".conv.weight"], sd[k + ".conv.bias"])
            out[("disp", i)] = torch.sigmoid(F.interpolate(raw, scale_factor=2, mode="bilinear"))
    return out


def motion_decoder(feats, ego_motion, sd, scales, out_dim):
    """networks/motion_decoder.py:34-91.  feats = [input, f/2, f/4, f/8, f/16, f/32]; ego (B,6,1,1)."""
    n_levels = len(feats)
    field = F.conv2d(100 * ego_motion, sd["_residual_translation.weight"], sd["_residual_translation.bias"])
    per_level = []
    for ii in range(n_levels):
        feat = feats[-1 - ii]
        up = F.interpolate(field, size=feat.shape[-2:], mode="bilinear", align_corners=False)
        x = torch.cat([up, feat], 1)
        x1 = F.conv2d(x, sd[f"refine_motion_conv{ii}.0.weight"], sd[f"refine_motion_conv{ii}.0.bias"], padding=1)
        x2 = F.conv2d(x1, sd[f"refine_motion_conv{ii}.1.weight"], sd[f"refine_motion_conv{ii}.1.bias"], padding=1)
        red = F.conv2d(torch.cat([x1, x2], 1), sd[f"refine_motion_redu{ii}.weight"], sd[f"refine_motion_redu{ii}.bias"])
        field = red + up
        per_level.append(field)
    out = {}
    for s in scales:
        raw = 0.01 * per_level[n_levels - 1 - s]
        if out_dim == 1:
            out[("motion_prob", s)] = raw
            out[("motion_mask", s)] = torch.sigmoid(raw)
        else:
            out[("complete_flow", s)] = raw
    return out


def pose_decoder(last_feature, sd, num_frames=2):
    """networks/pose_decoder.py:25-44 with one input feature list."""
    x = F.relu(F.conv2d(last_feature, sd["squeeze.weight"], sd["squeeze.bias"]))
    x = F.relu(F.conv2d(x, sd["pose0.weight"], sd["pose0.bias"], padding=1))
    x = F.relu(F.conv2d(x, sd["pose1.weight"], sd["pose1.bias"], padding=1))
    x = F.conv2d(x, sd["pose2.weight"], sd["pose2.bias"])
    out = 0.01 * x.mean(3).mean(2).view(-1, num_frames, 1, 6)
    return out[..., :3], out[..., 3:]


def transformation_from_parameters(axisangle, translation, invert=False):
    """networks/layers.py:7-82 (Rodrigues with angle+1e-7; invert -> R^T @ Trans(-t))."""
    B = axisangle.shape[0]
    v = axisangle.reshape(B, 3)
    angle = v.norm(dim=1, keepdim=True)
    axis = v / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[:, 0:1], axis[:, 1:2], axis[:, 2:3]
    zero, one = torch.zeros_like(ca), torch.ones_like(ca)
    R = torch.cat([x * x * C + ca, x * y * C - z * sa, z * x * C + y * sa, zero,
                   x * y * C + z * sa, y * y * C + ca, y * z * C - x * sa, zero,
                   z * x * C - y * sa, y * z * C + x * sa, z * z * C + ca, zero,
                   zero, zero, zero, one], 1).reshape(B, 4, 4)
    t = translation.reshape(B, 3)
    if invert:
        R = R.transpose(1, 2)
        t = -t
    T = torch.eye(4, dtype=v.dtype).repeat(B, 1, 1)
    T = torch.cat([T[:, :, :3], torch.cat([t, torch.ones(B, 1, dtype=v.dtype)], 1).unsqueeze(2)], 2)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


# ------------------------------------------------------------------------------------------------
# model wiring (networks/model.py:58-149) and one optimisation step (Trainer.py:140-151)
# ------------------------------------------------------------------------------------------------


class OracleModel(torch.nn.Module):
    """Encoders are injected modules; decoder weights are ParameterDicts with the reference's keys
    ('.' replaced by '/' because ParameterDict forbids dots)."""

    def __init__(self, depth_model, scales, frame_ids, depth_enc, pose_enc, motion_enc, decoder_states):
        super().__init__()
        self.depth_model, self.scales, self.frame_ids = depth_model, list(scales), list(frame_ids)
        self.depth_enc, self.pose_enc, self.motion_enc = depth_enc, pose_enc, motion_enc
        self.dec = torch.nn.ModuleDict()
        for name in ("depth_dec", "pose_dec", "motion_dec", "motion_mask"):
            self.dec[name] = torch.nn.ParameterDict({k.replace(".", "/"): torch.nn.Parameter(v.clone().float())
                                                     for k, v in decoder_states[name].items()})
        self.bool_CmpFlow, self.bool_MotMask = True, True
        self.network2modules = {"Depth": [self.depth_enc, self.dec["depth_dec"]], "Pose": [self.pose_enc, self.dec["pose_dec"]],
                                "CmpFlow": [self.motion_enc, self.dec["motion_dec"]],
                                "MotMask": [self.motion_enc, self.dec["motion_mask"]]}

    def _sd(self, name):
        return {k.replace("/", "."): v for k, v in self.dec[name].items()}

    def parameters_by_names(self, names):
        seen, out = set(), []
        for n in names:
            for m in self.network2modules[n]:
                for p in m.parameters():
                    if id(p) not in seen:
                        seen.add(id(p))
                        out.append(p)
        return out

    def forward(self, inputs):
        outputs = {}
        dec = depth_decoder_md2 if self.depth_model == "monodepthv2" else depth_decoder_lite
        for f in self.frame_ids:
            for (name, s), v in dec(self.depth_enc(inputs[("color_aug", f, 0)]), self._sd("depth_dec"), self.scales).items():
                outputs[(name, f, s)] = v
        for f in self.frame_ids[1:]:
            feats = self.pose_enc(torch.cat([inputs[("color_aug", f, 0)], inputs[("color_aug", 0, 0)]], 1))
            aa, tr = pose_decoder(feats[-1], self._sd("pose_dec"))
            aa, tr = aa[:, 0], tr[:, 0]
            outputs[("axisangle", 0, f)], outputs[("translation", 0, f)] = aa, tr
            outputs[("cam_T_cam", 0, f)] = transformation_from_parameters(aa, tr, invert=True)
        if self.bool_CmpFlow or self.bool_MotMask:
            trip = torch.cat([inputs[("color_aug", -1, 0)], inputs[("color_aug", 0, 0)], inputs[("color_aug", 1, 0)]], 1)
            feats = [trip] + list(self.motion_enc(trip))
            ego_t = (outputs[("translation", 0, -1)].detach() - outputs[("translation", 0, 1)].detach()) / 2
            ego_a = (outputs[("axisangle", 0, -1)].detach() - outputs[("axisangle", 0, 1)].detach()) / 2
            ego = torch.cat((ego_t, ego_a), -1).permute(0, 2, 1).unsqueeze(3)
            if self.bool_CmpFlow:
                for (name, s), v in motion_decoder(feats, ego, self._sd("motion_dec"), self.scales, 3).items():
                    outputs[(name, -1, s)], outputs[(name, 1, s)] = -1 * v, 1 * v
            if self.bool_MotMask:
                for (name, s), v in motion_decoder(feats, ego, self._sd("motion_mask"), self.scales, 1).items():
                    outputs[(name, -1, s)], outputs[(name, 1, s)] = v, v
        return outputs


class OracleTrainer:
    """process_batch + backward + Adam on CPU (Trainer.py:140-151,197-211,466-497)."""

    def __init__(self, model, height, width, learning_rate=1e-4, **loss_kw):
        self.model = model
        self.cfg = vs.LossConfig(height, width, model.scales, frame_ids=model.frame_ids, **loss_kw)
        self.lr = learning_rate
        self.step, self.steps_per_epoch = 0, 100
        self.optimizer = None
        # host RNG of the RANSAC ground prior (tools.py:126); replace for reproducible runs
        self.rand_index_fn = lambda n, k: np.random.choice(np.arange(n), k, replace=True)

    def setup_phase(self, phase):
        self.cfg.set_phase(phase)
        self.model.bool_CmpFlow, self.model.bool_MotMask = self.cfg.bool_CmpFlow, self.cfg.bool_MotMask
        factor = 0.5 if phase == "fine_tune" else 1.0
        self.optimizer = torch.optim.Adam(self.model.parameters_by_names(self.cfg.network_names), self.lr * factor)

    def process_batch(self, inputs, noise=None):
        outputs = self.model(inputs)
        vs.generate_images_pred(self.cfg, inputs, outputs)
        if noise is None and self.cfg.automask:
            B = inputs[("color", 0, 0)].shape[0]
            noise = {s: torch.randn(B, len(self.cfg.frame_ids) - 1, self.cfg.height, self.cfg.width) for s in self.cfg.scales}
        ground_fn = None
        if self.cfg.g["d_ground"] > 0 and self.cfg.bool_MotMask and "Depth" in self.cfg.network_names:
            from .ground import make_ground_fn
            ground_fn = make_ground_fn(self.cfg, self.rand_index_fn)
        losses = vs.compute_losses(self.cfg, inputs, outputs, self.step, self.steps_per_epoch, noise=noise, ground_fn=ground_fn)
        return outputs, losses

    def train_step(self, inputs, noise=None):
        outputs, losses = self.process_batch(inputs, noise)
        self.optimizer.zero_grad()
        losses["loss"].backward()
        self.optimizer.step()
        self.step += 1
        return outputs, losses
