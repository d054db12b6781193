from .det.dbnet import DBNetPP

__all__ = ["DBNetPP"]
