"""B200-native OCR inference path (drop-in for ZenHKD/ocr-vi-invoice `model/det`, and for the callers around it).

    from ocr_vi_invoice_b200.model.det.dbnet import DBNetPP        # reference: model/det/dbnet.py
    from ocr_vi_invoice_b200.postprocess import DBPostProcessor    # reference: src/det/test.py:46
    from ocr_vi_invoice_b200.preprocess import resize_image_for_det  # reference: src/pipeline/pipeline2.py:34
    from ocr_vi_invoice_b200.recprep import crop_image, preprocess_for_recognition  # test.py:123, pipeline2.py:92
    from ocr_vi_invoice_b200.model.rec2.svtrv2 import SVTRv2       # reference: model/rec2/svtrv2.py:409
    from ocr_vi_invoice_b200.pipeline import Detector, OCRPipeline  # the page loop of pipeline2.py:279-352

Everything numerical runs in hand-written sm_100a CUDA (csrc/, C-ABI in include/dbnet_b200.h); PyTorch only
provides device memory, streams and CUDA-graph capture.  There is no CPU / eager fallback.
"""
__version__ = "0.2.0"
