"""FPN + adaptive-scale-fusion parameter tree mirroring model/det/neck.py:6-79.

lateral 1x1 convs (with bias, no BN, neck.py:16), 3x3 smoothing ConvBnRelu (neck.py:21) and the ASF attention
conv (1024 -> 4, neck.py:55).  The top-down nearest upsample + add is fused into the lateral conv's epilogue and
the whole ScaleFeatureSelection.forward is one fused kernel (csrc/elementwise.cu, asf_kernel).
"""
import torch.nn as nn

from .layers import ConvBnRelu, _no_eager


class ScaleFeatureSelection(nn.Module):
    def __init__(self, num_levels, inner_channels):
        super().__init__()
        if num_levels != 4 or inner_channels != 256:
            raise NotImplementedError("B200 path: ASF kernel is specialised for 4 levels x 256 channels")
        self.inner_channels = inner_channels
        self.conv_atten = nn.Conv2d(inner_channels * num_levels, num_levels, 1)

    forward = _no_eager


class FPN_ASF(nn.Module):
    def __init__(self, in_channels_list, inner_channels=256, use_asf=True):
        super().__init__()
        if not use_asf:
            raise NotImplementedError("B200 path: DBNet++ always uses ASF (neck.py:43-46 fallback not built)")
        self.use_asf = use_asf
        self.inner_channels = inner_channels
        self.in_channels_list = list(in_channels_list)
        self.lateral_convs = nn.ModuleList(nn.Conv2d(c, inner_channels, 1) for c in in_channels_list)
        self.fpn_convs = nn.ModuleList(ConvBnRelu(inner_channels, inner_channels, 3, 1, 1) for _ in in_channels_list)
        self.asf = ScaleFeatureSelection(len(self.in_channels_list), inner_channels)

    forward = _no_eager
