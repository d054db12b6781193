"""Parameter holder mirroring the reference's DeformableConv2d (model/det/dcn.py:5-59).

`offset_mask_conv` (3x3, Cin -> 27, bias, zero-initialised as dcn.py:28-29) and `weight` keep their names; the
forward -- offset conv, sigmoid mask, bilinear sampling and the GEMM of torchvision.ops.deform_conv2d
(dcn.py:41-57) -- is two launches of the tcgen05 implicit-GEMM kernel (csrc/igemm.cu, modes CONV and DCN).
"""
import torch
import torch.nn as nn

from .layers import _no_eager


class DeformableConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, dilation=1, bias=False):
        super().__init__()
        if kernel_size != 3 or padding != 1 or dilation != 1:
            raise NotImplementedError("B200 path: DCN is 3x3, padding 1, dilation 1 (all DBNet++ uses)")
        self.kernel_size, self.stride, self.padding, self.dilation = kernel_size, stride, padding, dilation
        self.in_channels, self.out_channels = in_channels, out_channels
        self.offset_mask_conv = nn.Conv2d(in_channels, 3 * kernel_size * kernel_size, kernel_size=kernel_size,
                                          stride=stride, padding=padding, dilation=dilation, bias=True)
        nn.init.constant_(self.offset_mask_conv.weight, 0)
        nn.init.constant_(self.offset_mask_conv.bias, 0)
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=1)
        if self.bias is not None:
            nn.init.constant_(self.bias, 0)

    forward = _no_eager
