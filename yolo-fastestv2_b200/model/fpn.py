"""Parameter containers for LightFPN and its DWConvblock heads (mirror of reference model/fpn.py:5-64:
conv1x1_{2,3}.{0,1}, {cls,reg}_head_{2,3}.block.{0,1,3,4,5,6,8,9}).  Compute: csrc/k_tcnet.cu (tc_pw_kernel: FPN reducers; tc_head*_kernel: DWConvblock heads + folded output convs)."""
import torch.nn as nn

from model.backbone.shufflenetv2 import _WeightsOnly, _conv_bn


class DWConvblock(_WeightsOnly):
    def __init__(self, input_channels, output_channels, size):
        super().__init__()
        self.size, self.input_channels, self.output_channels = size, input_channels, output_channels
        c = output_channels      # the reference ignores input_channels: every conv is c -> c (fpn.py:12-24)

        def dw():
            return [nn.Conv2d(c, c, size, 1, 2, groups=c, bias=False), nn.BatchNorm2d(c), nn.ReLU(inplace=True)]

        self.block = nn.Sequential(*(dw() + _conv_bn(c, c, 1, 1) + dw() + _conv_bn(c, c, 1, 1)))


class LightFPN(_WeightsOnly):
    def __init__(self, input2_depth, input3_depth, out_depth):
        super().__init__()
        self.conv1x1_2 = nn.Sequential(*_conv_bn(input2_depth, out_depth, 1, 1, relu=True))
        self.conv1x1_3 = nn.Sequential(*_conv_bn(input3_depth, out_depth, 1, 1, relu=True))
        self.cls_head_2 = DWConvblock(input2_depth, out_depth, 5)
        self.reg_head_2 = DWConvblock(input2_depth, out_depth, 5)
        self.reg_head_3 = DWConvblock(input3_depth, out_depth, 5)
        self.cls_head_3 = DWConvblock(input3_depth, out_depth, 5)
