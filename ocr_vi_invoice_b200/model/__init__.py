from .det.dbnet import DBNetPP
from .rec2.svtrv2 import SVTRv2

__all__ = ["DBNetPP", "SVTRv2"]
