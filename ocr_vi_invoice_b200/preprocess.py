"""GPU side of the detection pre-processing (reference: resize_image_for_det, src/pipeline/pipeline2.py:34-40, and the
host normalisation at pipeline2.py:312-314).

`resize_image_for_det(image, image_size)` keeps the reference's name, arguments and return value
`(resized, (scale_h, scale_w))`, but `image` is a uint8 HWC CUDA tensor and `resized` a uint8 HWC CUDA tensor produced by
the `b200_resize_u8_linear` kernel (bit-identical to cv2.resize).  The normalisation + HWC->CHW + bf16 packing is fused
into the stem packing kernel (`DBNetPP.forward` accepts the uint8 NHWC batch directly), so a decoded image goes
host -> device once, as bytes.  No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import ops


def det_size(h: int, w: int, image_size: int = 640) -> Tuple[Tuple[int, int], Tuple[float, float]]:
    """(new_h, new_w) rounded to multiples of 32 and the (scale_h, scale_w) the caller divides boxes by."""
    scale = image_size / max(h, w)
    new_h = int(np.round(h * scale / 32) * 32)
    new_w = int(np.round(w * scale / 32) * 32)
    return (new_h, new_w), (new_h / h, new_w / w)


def resize_image_for_det(image: torch.Tensor, image_size: int = 640, out: torch.Tensor = None):
    if not (isinstance(image, torch.Tensor) and image.is_cuda and image.dtype == torch.uint8 and image.dim() == 3
            and image.shape[2] == 3):
        raise RuntimeError("resize_image_for_det (B200): expects a uint8 HWC CUDA tensor (no CPU fallback)")
    image = image.contiguous()
    (new_h, new_w), scales = det_size(image.shape[0], image.shape[1], image_size)
    if new_h <= 0 or new_w <= 0:
        raise ValueError("resize_image_for_det: image too thin for a multiple-of-32 target")
    if out is None:
        out = torch.empty(new_h, new_w, 3, dtype=torch.uint8, device=image.device)
    elif tuple(out.shape) != (new_h, new_w, 3):
        raise ValueError(f"out must be ({new_h}, {new_w}, 3)")
    with torch.cuda.device(image.device):
        ops.resize_u8_linear(image, out)
    return out, scales
