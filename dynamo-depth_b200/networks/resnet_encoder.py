"""ResNet encoders (reference: networks/resnet_encoder.py).  Kept in PyTorch/cuDNN on purpose: the
north-star reserves the tensor cores for these contractions and hand-writes only the decoders and the
loss path.  state_dict keys are torchvision's under the attribute `encoder` (including the unused
`encoder.fc.*`), so reference checkpoints load unchanged."""
import numpy as np
import torch
import torch.nn as nn
import torchvision.models as tvm

_FACTORIES = {18: (tvm.resnet18, "ResNet18_Weights"), 34: (tvm.resnet34, "ResNet34_Weights"),
              50: (tvm.resnet50, "ResNet50_Weights"), 101: (tvm.resnet101, "ResNet101_Weights"),
              152: (tvm.resnet152, "ResNet152_Weights")}


def _imagenet_state(num_layers):
    factory, weights_name = _FACTORIES[num_layers]
    weights = getattr(tvm, weights_name).IMAGENET1K_V1
    return weights.get_state_dict(progress=False)   # needs network access or a populated torch hub cache


def resnet_multiimage_input(num_layers, pretrained=False, num_input_images=1, inp_disp=False):
    """ResNet-18/50 whose stem takes `num_input_images` stacked frames (resnet_encoder.py:64-92)."""
    assert num_layers in (18, 50), "Can only run with 18 or 50 layer resnet"
    per_img = 4 if inp_disp else 3
    net = _FACTORIES[num_layers][0](weights=None)
    stem = nn.Conv2d(num_input_images * per_img, 64, kernel_size=7, stride=2, padding=3, bias=False)
    nn.init.kaiming_normal_(stem.weight, mode="fan_out", nonlinearity="relu")
    net.conv1 = stem
    if pretrained:
        state = _imagenet_state(num_layers)
        w = nn.init.kaiming_normal_(torch.ones(64, per_img * num_input_images, 7, 7))
        for k in range(num_input_images):   # tile the RGB stem over the frames, keep the average response
            w[:, per_img * k:per_img * k + 3] = state["conv1.weight"] / num_input_images
        state["conv1.weight"] = w
        net.load_state_dict(state)
    return net


class ResnetEncoder(nn.Module):
    def __init__(self, num_layers, pretrained, num_input_images=1, inp_disp=False):
        super().__init__()
        if num_layers not in _FACTORIES:
            raise ValueError(f"{num_layers} is not a valid number of resnet layers")
        self.num_ch_enc = np.array([64, 64, 128, 256, 512])
        if num_input_images == 1:
            assert not inp_disp, "single input image cannot be RGBD"
            self.encoder = _FACTORIES[num_layers][0](weights=None)
            if pretrained:
                self.encoder.load_state_dict(_imagenet_state(num_layers))
        else:
            self.encoder = resnet_multiimage_input(num_layers, pretrained, num_input_images, inp_disp)
        if num_layers > 34:
            self.num_ch_enc[1:] *= 4

    # NHWC activations + weights on CUDA: cuDNN then runs its native tensor-core kernels without the per-layer
    # NCHW<->NHWC conversions and with the faster NHWC batch-norm kernels (same arithmetic, ~7 % of the bs32 step).
    channels_last = True
    # BN + ReLU (+ identity) of the BasicBlocks through dd_bn_act_nhwc_*: measured SLOWER than cuDNN's single-read NHWC batch
    # norm + ATen's ReLU / add on the B200 (bs32 step 114.5 -> 117.0 ms), so opt-in only (tests exercise it)
    fused_blocks = False

    def to_channels_last(self):
        """Convert the trunk's weights once.  Must happen BEFORE a gradient arena / optimiser is built over the parameters:
        Module.to(memory_format=...) re-creates `.grad` tensors too, which would silently detach them from the arena."""
        if self.channels_last and not getattr(self, "_cl_ready", False):
            self.encoder.to(memory_format=torch.channels_last)
            self._cl_ready = True

    @staticmethod
    def _basic_block(blk, x):
        """torchvision BasicBlock.forward with BN + ReLU and BN + identity + ReLU as fused channels_last kernels
        (dd_bn_act_nhwc_*; eval mode and exotic channel counts fall back to the torch ops inside bn_act_nhwc)."""
        from dd_b200 import functional as DF
        out = DF.bn_act_nhwc(blk.conv1(x), blk.bn1, "relu")
        identity = x if blk.downsample is None else DF.bn_act_nhwc(blk.downsample[0](x), blk.downsample[1], "none")
        return DF.bn_act_nhwc(blk.conv2(out), blk.bn2, "relu", residual=identity)

    def _run_layer(self, layer, x, fused):
        for blk in layer:
            plain_ds = blk.downsample is None or (len(blk.downsample) == 2 and isinstance(blk.downsample[1], nn.BatchNorm2d))
            x = self._basic_block(blk, x) if (fused and isinstance(blk, tvm.resnet.BasicBlock) and plain_ds) else blk(x)
        return x

    def forward(self, input_image):
        e = self.encoder
        cl = self.channels_last and input_image.is_cuda
        fused = cl and self.fused_blocks
        if cl:
            from dd_b200 import functional as DF
            self.to_channels_last()
            input_image = input_image.contiguous(memory_format=torch.channels_last)
        x = e.conv1((input_image - 0.45) / 0.225)
        x = DF.bn_act_nhwc(x, e.bn1, "relu") if fused else e.relu(e.bn1(x))
        self.features = [x]
        if cl and x.shape[1] % 4 == 0 and (e.maxpool.kernel_size, e.maxpool.stride, e.maxpool.padding) == (3, 2, 1):
            x = DF.maxpool3x3s2(x)   # gather-style NHWC max-pool (csrc/pool.cu)
        else:
            x = e.maxpool(x)
        x = self._run_layer(e.layer1, x, fused)
        self.features.append(x)
        for layer in (e.layer2, e.layer3, e.layer4):
            x = self._run_layer(layer, x, fused)
            self.features.append(x)
        if cl:
            # the decoders' kernels read NCHW: one tiled transpose per feature map here (and one for its gradient) instead of a
            # generic strided copy inside every consumer
            self.features = [DF.to_nchw(f) for f in self.features]
        return self.features
