"""DB head parameter tree mirroring model/det/head.py:5-48 (two identical branches, indices 0,1,2,4 hold params)."""
import torch.nn as nn

from .layers import ConvBnRelu, _no_eager


class DBHead(nn.Module):
    def __init__(self, in_channels, k=50):
        super().__init__()
        if in_channels != 256:
            raise NotImplementedError("B200 path: head kernels are specialised for 256 -> 64 -> 64 -> 1 channels")
        self.k = k

        def branch():
            return nn.Sequential(
                ConvBnRelu(in_channels, in_channels // 4, 3, 1, 1),
                nn.ConvTranspose2d(in_channels // 4, in_channels // 4, 2, 2),
                nn.BatchNorm2d(in_channels // 4),
                nn.ReLU(inplace=True),
                nn.ConvTranspose2d(in_channels // 4, 1, 2, 2),
            )

        self.bin_conv = branch()
        self.thresh_conv = branch()

    forward = _no_eager
