"""Model wiring (reference: networks/model.py): depth / pose / scene-flow / motion-mask networks, phase
flags, forward dict contract `(name, frame_id, scale)`, per-module checkpoints.  Encoders run in
PyTorch/cuDNN, every decoder in the hand-written kernels."""
import os
import os.path as osp

import torch
import torch.nn as nn

from .depth_decoder import DepthDecoder, LiteDepthDecoder
from .depth_encoder import LiteMono
from .layers import transformation_from_parameters
from .motion_decoder import MotionDecoder
from .pose_decoder import PoseDecoder
from .resnet_encoder import ResnetEncoder


MODULE_ORDER = ("depth_enc", "depth_dec", "pose_enc", "pose_dec", "motion_enc", "motion_dec", "motion_mask")


class Model(nn.Module):
    network2modules = None   # filled per instance (kept as an attribute like the reference)

    def __init__(self, options):
        super().__init__()
        self.opt = options
        pre = self.opt.weights_init == "pretrained"
        if self.opt.depth_model == "monodepthv2":
            self.depth_enc = ResnetEncoder(self.opt.encoder_num_layers, pre)
            self.depth_dec = DepthDecoder(self.depth_enc.num_ch_enc, self.opt.scales)
        elif self.opt.depth_model == "litemono":
            self.depth_enc = LiteMono(model="lite-mono-8m", drop_path_rate=0.4, pretrained=pre)
            self.depth_dec = LiteDepthDecoder(self.depth_enc.num_ch_enc, self.opt.scales)
        else:
            raise Exception(f"Model Name {self.opt.depth_model} not recognized.")
        self.pose_enc = ResnetEncoder(self.opt.encoder_num_layers, pre, num_input_images=2, inp_disp=False)
        self.pose_dec = PoseDecoder(self.pose_enc.num_ch_enc, num_input_features=1, num_frames_to_predict_for=2)
        self.motion_enc = ResnetEncoder(self.opt.encoder_num_layers, pre, num_input_images=3, inp_disp=False)
        self.motion_dec = MotionDecoder(self.pose_enc.num_ch_enc, self.opt.scales, num_input_images=3, inp_disp=False, out_dim=3)
        self.motion_mask = MotionDecoder(self.pose_enc.num_ch_enc, self.opt.scales, num_input_images=3, inp_disp=False, out_dim=1)
        self.network2modules = {"Depth": ["depth_enc", "depth_dec"], "Pose": ["pose_enc", "pose_dec"],
                                "CmpFlow": ["motion_enc", "motion_dec"], "MotMask": ["motion_enc", "motion_mask"]}
        # Fixed order (the reference's list(set(...)) depends on PYTHONHASHSEED): the gradient-arena layout, the Adam
        # state order of a checkpoint and the construction-time broadcast must be identical in every process.
        self.module_names = [m for m in MODULE_ORDER if any(m in mods for mods in self.network2modules.values())]
        self.bool_CmpFlow = True
        self.bool_MotMask = True
        # opt-in: skip the depth passes on frames -1/+1, which no loss term consumes.
        # Off by default because it changes BatchNorm running statistics / the DropPath RNG stream.
        self.skip_unused_depth = bool(getattr(options, "skip_unused_depth", False))
        # opt-in: the two pose-encoder calls (reference model.py:82-86) as ONE batch of 2B image pairs.
        # Off by default: the pose encoder's BatchNorm then normalises over both pairs together (train mode).
        self.batch_pose_pairs = bool(getattr(options, "batch_pose_pairs", False))
        # Lite-Mono encoder linear layers: hand-written tcgen05 3xTF32 kernel (default, fp32 accuracy), torch fp32 SIMT, or
        # (opt-in) cuBLAS single-pass TF32
        from . import depth_encoder as _de
        _de.EncoderLinear.tf32 = bool(getattr(options, "encoder_tf32_linear", False))
        _de.EncoderLinear.mode = getattr(options, "encoder_linear", "tc3x")

    def forward(self, inputs):
        outputs = {}
        self.predict_depths(inputs, outputs)
        self.predict_poses(inputs, outputs)
        self.predict_motions(inputs, outputs)
        return outputs

    def predict_depths(self, inputs, outputs):
        frames = self.opt.frame_ids[:1] if self.skip_unused_depth else self.opt.frame_ids
        for f in frames:
            disp = self.depth_dec(self.depth_enc(inputs["color_aug", f, 0]))
            for (name, s), v in disp.items():
                outputs[(name, f, s)] = v

    def predict_poses(self, inputs, outputs):
        frames = self.opt.frame_ids[1:]
        pairs = [torch.cat([inputs["color_aug", f, 0], inputs["color_aug", 0, 0]], 1) for f in frames]   # target frame always last
        batched = None
        if self.batch_pose_pairs and len(frames) > 1:
            B = pairs[0].shape[0]
            feats_all = self.pose_enc(torch.cat(pairs, 0))
            aa_all, tr_all = self.pose_dec([feats_all])
            batched = [([ft[i * B:(i + 1) * B] for ft in feats_all], aa_all[i * B:(i + 1) * B], tr_all[i * B:(i + 1) * B])
                       for i in range(len(frames))]
        for i, f in enumerate(frames):
            pair = pairs[i]
            if batched is not None:
                feats, axisangle, translation = batched[i]
            else:
                feats = self.pose_enc(pair)
                axisangle, translation = self.pose_dec([feats])
            axisangle, translation = axisangle[:, 0], translation[:, 0]
            outputs[("pose_feats", 0, f)] = [pair] + feats
            outputs[("axisangle", 0, f)] = axisangle
            outputs[("translation", 0, f)] = translation
            outputs[("cam_T_cam", 0, f)] = transformation_from_parameters(axisangle, translation, invert=True)

    def predict_motion_feat(self, inputs, outputs):
        for gap in set(abs(f) for f in self.opt.frame_ids[1:]):
            triple = torch.cat([inputs["color_aug", -gap, 0], inputs["color_aug", 0, 0], inputs["color_aug", gap, 0]], 1)
            outputs[("motion_feats", 0, gap)] = [triple] + self.motion_enc(triple)

    def predict_motions(self, inputs, outputs):
        if not self.bool_CmpFlow and not self.bool_MotMask:
            return
        self.predict_motion_feat(inputs, outputs)
        for gap in set(abs(f) for f in self.opt.frame_ids[1:]):
            prev, nxt = -gap, gap
            feats = outputs[("motion_feats", 0, gap)]
            ego_t = (outputs[("translation", 0, prev)].detach() - outputs[("translation", 0, nxt)].detach()) / 2
            ego_a = (outputs[("axisangle", 0, prev)].detach() - outputs[("axisangle", 0, nxt)].detach()) / 2
            ego = torch.cat((ego_t, ego_a), -1).permute(0, 2, 1).unsqueeze(3)   # (B,6,1,1)
            if self.bool_CmpFlow:
                for (name, s), v in self.motion_dec(feats, ego).items():
                    outputs[(name, prev, s)] = -1 * v     # flow towards the past is the negated field
                    outputs[(name, nxt, s)] = 1 * v
            if self.bool_MotMask:
                for (name, s), v in self.motion_mask(feats, ego).items():
                    outputs[(name, prev, s)] = v
                    outputs[(name, nxt, s)] = v

    def prepare_memory_format(self):
        """One-time layout conversions of the PyTorch encoders (NHWC ResNet trunks on CUDA); call after .to(device) and
        before optimisers / gradient arenas are built."""
        for name in self.module_names:
            mod = getattr(self, name)
            if isinstance(mod, ResnetEncoder) and next(mod.parameters()).is_cuda:
                mod.to_channels_last()

    def modules_by_names(self, network_names):
        """Sub-module names of the given networks in the fixed MODULE_ORDER, each once (motion_enc is shared)."""
        wanted = set(m for n in network_names for m in self.network2modules[n])
        return [m for m in MODULE_ORDER if m in wanted]

    def parameters_by_names(self, network_names):
        params = []
        for m in self.modules_by_names(network_names):
            params += list(getattr(self, m).parameters())
        return params

    def named_parameters_by_names(self, network_names):
        """[("<module>.<param>", parameter)] in the order of parameters_by_names (layout check / checkpoints)."""
        named = []
        for m in self.modules_by_names(network_names):
            named += [(f"{m}.{k}", p) for k, p in getattr(self, m).named_parameters()]
        return named

    def save(self, save_folder):
        for name in self.module_names:
            state = getattr(self, name).state_dict()
            if "enc" in name:
                state["height"], state["width"] = self.opt.height, self.opt.width
            torch.save(state, osp.join(save_folder, f"{name}.pth"))

    def load(self, dev="cpu", verbose=True):
        folder = osp.expanduser(self.opt.load_ckpt)
        self.opt.load_ckpt = folder
        if not osp.isdir(folder):
            raise Exception(f"Cannot find folder {folder} (checkpoint download is not available offline)")
        for name in self.module_names:
            path = osp.join(folder, f"{name}.pth")
            if not osp.exists(path):
                if verbose:
                    print(f"|- Loading {name} weights... FAILED :: Path {path} not found")
                continue
            ckpt = torch.load(path, map_location=dev)
            if "height" in ckpt:
                if verbose and (ckpt["height"], ckpt["width"]) != (self.opt.height, self.opt.width):
                    print(f"|- === WARNING: self.opt ({self.opt.height},{self.opt.width}) != loaded ({ckpt['height']},{ckpt['width']})")
                ckpt.pop("height"), ckpt.pop("width")
            module = getattr(self, name)
            try:
                module.load_state_dict(ckpt)
            except Exception:
                own = module.state_dict()
                own.update({k: v for k, v in ckpt.items() if k in own and v.shape == own[k].shape})
                module.load_state_dict(own)

    def set_train(self):
        for name in self.module_names:
            getattr(self, name).train()

    def set_eval(self):
        for name in self.module_names:
            getattr(self, name).eval()
