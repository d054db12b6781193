from .dbnet import DBNetPP
from .backbone import ResNet
from .neck import FPN_ASF, ScaleFeatureSelection
from .head import DBHead
from .dcn import DeformableConv2d
from .layers import ConvBnRelu

__all__ = ["DBNetPP", "ResNet", "FPN_ASF", "ScaleFeatureSelection", "DBHead", "DeformableConv2d", "ConvBnRelu"]
