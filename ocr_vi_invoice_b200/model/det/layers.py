"""Parameter holder mirroring the reference's ConvBnRelu (model/det/layers.py:4-18).

Same sub-module names (`conv`, `bn`) so state_dict keys match; the arithmetic (conv + folded BN + ReLU) is
one fused tcgen05 implicit-GEMM launch planned by engine.py, so there is no forward here.
"""
import torch.nn as nn


def _no_eager(self, *args, **kwargs):
    raise RuntimeError(
        f"{type(self).__name__} is a parameter container of the B200 DBNet++ path; it has no eager forward. "
        "Call DBNetPP.forward on a CUDA tensor (there is no CPU / torch fallback).")


class ConvBnRelu(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, dilation=1, groups=1,
                 bias=False, activation=True):
        super().__init__()
        if dilation != 1 or groups != 1:
            raise NotImplementedError("B200 path: dilation/groups other than 1 are not used by DBNet++")
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                              bias=bias)
        self.bn = nn.BatchNorm2d(out_channels)
        self.activation = activation
        if activation:
            self.relu = nn.ReLU(inplace=True)

    forward = _no_eager
