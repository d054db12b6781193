"""ResNet-50(-DCN) parameter tree mirroring the reference's backbone wrapper (model/det/backbone.py:8-60).

The reference wraps torchvision.models.resnet50 and registers it twice (`self.model` and `self.layer1..4`,
backbone.py:34-37), which duplicates every backbone key in the state_dict and keeps the unused `fc`.  This
module rebuilds the same tree from plain nn.Conv2d / nn.BatchNorm2d holders (no torchvision import) with the
same names, registration order and initialisation, so `load_state_dict(strict=True)` of a reference
checkpoint works.  Compute is planned by engine.py.
"""
import torch.nn as nn

from .dcn import DeformableConv2d
from .layers import _no_eager


class Bottleneck(nn.Module):
    """torchvision Bottleneck v1.5 (stride on conv2), names as torchvision/models/resnet.py:86-163."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    forward = _no_eager


class _ResNet50(nn.Module):
    """Holder with torchvision.models.resnet50's attribute names (conv1, bn1, relu, maxpool, layer1-4, avgpool, fc)."""

    def __init__(self):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(64, 3, 1)
        self.layer2 = self._make_layer(128, 4, 2)
        self.layer3 = self._make_layer(256, 6, 2)
        self.layer4 = self._make_layer(512, 3, 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(2048, 1000)  # unused by detection, present in the reference's state_dict
        for m in self.modules():  # torchvision's init (resnet.py:208-213)
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes * 4:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes * 4, 1, stride=stride, bias=False),
                                       nn.BatchNorm2d(planes * 4))
        layers = [Bottleneck(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * 4
        layers += [Bottleneck(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    forward = _no_eager


class ResNet(nn.Module):
    def __init__(self, name="resnet50", pretrained=True, in_channels=3, dcn=False):
        super().__init__()
        if name != "resnet50":
            # model/det/backbone.py:12-15 also offers resnet18; the B200 path accelerates the checkpointed
            # resnet50 configuration only and refuses the rest instead of running them unaccelerated.
            raise NotImplementedError(f"B200 DBNet++ path implements backbone 'resnet50' only (got {name!r})")
        if in_channels != 3:
            raise NotImplementedError("B200 DBNet++ path: the stem kernel is specialised for 3 input channels")
        if pretrained:
            # model/det/backbone.py:17 would download torchvision's ImageNet weights here.  The reference's own callers
            # build DBNetPP() with this default and then load a checkpoint (src/det/val.py:147-150), so the drop-in must
            # construct: warn, keep the random init (no network, nothing is downloaded) and expect load_state_dict.
            import warnings
            warnings.warn("B200 DBNet++ path: pretrained=True does not download ImageNet weights (no network access); "
                          "the backbone keeps its random init until load_state_dict() is called", RuntimeWarning,
                          stacklevel=3)
        self.model = _ResNet50()
        self.out_channels = [256, 512, 1024, 2048]
        self.dcn = bool(dcn)
        if dcn:  # c3, c4, c5 (backbone.py:27-31)
            self._apply_dcn(self.model.layer2)
            self._apply_dcn(self.model.layer3)
            self._apply_dcn(self.model.layer4)
        self.layer1 = nn.Sequential(self.model.conv1, self.model.bn1, self.model.relu, self.model.maxpool,
                                    self.model.layer1)
        self.layer2 = self.model.layer2
        self.layer3 = self.model.layer3
        self.layer4 = self.model.layer4

    def _apply_dcn(self, layer):
        for block in layer:
            c2 = block.conv2
            block.conv2 = DeformableConv2d(c2.in_channels, c2.out_channels, kernel_size=3, stride=c2.stride[0],
                                           padding=c2.padding[0], bias=False)

    forward = _no_eager
