"""Parameter containers for the ShuffleNetV2 backbone (mirror of the reference's
model/backbone/shufflenetv2.py:5-114 module tree, so state_dict keys/shapes are identical:
first_conv.{0,1}, stage{2,3,4}.{i}.branch_main.{0,1,3,4,5,6}, .branch_proj.{0,1,2,3}).

These modules own weights only.  The arithmetic runs in libyfv2.so (csrc/k_stem.cu: stem + max-pool; k_blk.cu / k_tail.cu: chained stride-1 blocks; k_tcnet.cu: stride-2 blocks),
dispatched from Detector.forward; calling a sub-module directly is not supported."""
import os

import torch
import torch.nn as nn


def _conv_bn(cin, cout, k, stride, groups=1, relu=False):
    layers = [nn.Conv2d(cin, cout, k, stride, k // 2, groups=groups, bias=False), nn.BatchNorm2d(cout)]
    if relu:
        layers.append(nn.ReLU(inplace=True))
    return layers


class _WeightsOnly(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("%s holds weights only; run the whole Detector (CUDA kernels in libyfv2.so)" % type(self).__name__)


class ShuffleV2Block(_WeightsOnly):
    def __init__(self, inp, oup, mid_channels, *, ksize, stride):
        super().__init__()
        if stride not in (1, 2):
            raise ValueError("stride must be 1 or 2")
        self.stride, self.inp, self.mid_channels, self.ksize, self.pad = stride, inp, mid_channels, ksize, ksize // 2
        # pw + BN + ReLU, dw + BN, pw + BN + ReLU
        self.branch_main = nn.Sequential(*(_conv_bn(inp, mid_channels, 1, 1, relu=True)
                                           + _conv_bn(mid_channels, mid_channels, ksize, stride, groups=mid_channels)
                                           + _conv_bn(mid_channels, oup - inp, 1, 1, relu=True)))
        # dw + BN, pw + BN + ReLU (downsampling blocks only)
        self.branch_proj = (nn.Sequential(*(_conv_bn(inp, inp, ksize, stride, groups=inp) + _conv_bn(inp, inp, 1, 1, relu=True)))
                            if stride == 2 else None)


class ShuffleNetV2(_WeightsOnly):
    stage_repeats = (4, 8, 4)

    def __init__(self, stage_out_channels, load_param):
        super().__init__()
        self.stage_out_channels = stage_out_channels
        cin = stage_out_channels[1]
        self.first_conv = nn.Sequential(*_conv_bn(3, cin, 3, 2, relu=True))
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        for si, rep in enumerate(self.stage_repeats):
            cout = stage_out_channels[si + 2]
            blocks = [ShuffleV2Block(cin, cout, mid_channels=cout // 2, ksize=3, stride=2)]
            blocks += [ShuffleV2Block(cout // 2, cout, mid_channels=cout // 2, ksize=3, stride=1) for _ in range(rep - 1)]
            setattr(self, "stage%d" % (si + 2), nn.Sequential(*blocks))
            cin = cout
        if load_param:
            print("load param...")
        else:
            self._initialize_weights()

    def _initialize_weights(self):
        # same contract as the reference: cwd-relative pretrained backbone, strict load
        print("initialize_weights...")
        path = "./model/backbone/backbone.pth"
        if not os.path.exists(path):
            raise FileNotFoundError("%s not found (the reference loads its pretrained backbone from the current "
                                    "working directory when load_param=False)" % path)
        self.load_state_dict(torch.load(path, map_location="cpu"), strict=True)
