"""B200-native DBNet++ text-detection inference path (drop-in for ZenHKD/ocr-vi-invoice `model/det`).

    from ocr_vi_invoice_b200.model.det.dbnet import DBNetPP      # reference: model/det/dbnet.py
    from ocr_vi_invoice_b200.postprocess import DBPostProcessor  # reference: src/det/test.py:46

Everything numerical runs in hand-written sm_100a CUDA (csrc/, C-ABI in include/dbnet_b200.h); PyTorch only
provides device memory and streams.  There is no CPU / eager fallback.
"""
__version__ = "0.1.0"
