"""GPU side of the detection -> recognition hand-over.

Reference functions mirrored here (same names, argument meaning and results, CUDA tensors instead of numpy arrays):
  * `crop_image(img, box)`                         src/det/test.py:123-130
  * `preprocess_for_recognition(crop, img_size)`   src/pipeline/pipeline2.py:92-128
  * the crop + batch-assembly loop of the pipeline pipeline2.py:323-328, 339-343, 149-159  -> `RecognitionCropper`

`RecognitionCropper.from_detector` consumes the DEVICE box records `DBPostProcessor.run_device` produced for a batch of
pages and writes the (N, 3, 32, 256) float32 recogniser input in one kernel launch, nothing synchronises and nothing
returns to the host; `fetch` is there for callers (and tests) that want the rectangles.  Results are bit-identical to the
reference's numpy / cv2 / torch code (oracle/recprep_oracle.py, pinned by tests/golden/recprep_reference.npz).  No CPU
fallback: CPU tensors raise.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops

_RECT_DTYPE = np.dtype([("image", "<i4"), ("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("status", "<i4")])
assert _RECT_DTYPE.itemsize == 24

STATUS_OK, STATUS_EMPTY, STATUS_DEGENERATE = _lib.REC_OK, _lib.REC_EMPTY, _lib.REC_DEGENERATE


class DegenerateCropError(ValueError):
    """A crop so tall and thin that int(w * 32 / h) == 0: the reference's cv2.resize raises cv2.error there
    (pipeline2.py:106, `inv_scale_x > 0`)."""


def _need_cuda_page(img) -> torch.Tensor:
    if not (isinstance(img, torch.Tensor) and img.is_cuda and img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3):
        raise RuntimeError("recognition crops (B200): expects a uint8 HWC CUDA tensor with 3 channels (no CPU fallback)")
    return img


def crop_image(img: torch.Tensor, box) -> torch.Tensor:
    """test.py:123-130 on a CUDA page: returns the VIEW img[y:y+bh, x:x+bw] (no kernel, no copy).
    `box` is a host (K, 2) integer array like the rescaled boxes of pipeline2.py:327."""
    _need_cuda_page(img)
    h, w = img.shape[:2]
    box = np.asarray(box)
    # cv2.boundingRect of an integer point set
    x, y = int(box[:, 0].min()), int(box[:, 1].min())
    bw, bh = int(box[:, 0].max()) - x + 1, int(box[:, 1].max()) - y + 1
    x = max(0, x)
    y = max(0, y)
    bw = min(bw, w - x)
    bh = min(bh, h - y)
    if bw <= 0 or bh <= 0:
        return img[0:0, 0:0]
    return img[y:y + bh, x:x + bw]


class RecognitionCropper:
    """Batched crop + recognition pre-processing with reusable device buffers (one shape cached)."""

    def __init__(self, img_size: Tuple[int, int] = (32, 256), capacity: int = 1024, device=None,
                 mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD):
        self.img_size = (int(img_size[0]), int(img_size[1]))
        self.capacity = int(capacity)
        self.device = device
        self.mean, self.std = tuple(mean), tuple(std)
        self._buf = None

    def _buffers(self, dev):
        if self._buf is None or self._buf["out"].device != dev:
            th, tw = self.img_size
            self._buf = {
                "out": torch.empty(self.capacity, 3, th, tw, dtype=torch.float32, device=dev),
                "rects": torch.empty(self.capacity * 6, dtype=torch.int32, device=dev),
                "total": torch.zeros(1, dtype=torch.int32, device=dev),
            }
        return self._buf

    # ------------------------------------------------------------------ device side, no synchronisation
    def from_detector(self, det_buf: dict, pages: Sequence[torch.Tensor], scales: Sequence[Tuple[float, float]],
                      max_candidates: int):
        """det_buf: the dict `DBPostProcessor.run_device` returned (device `boxes`, `counts`) for len(pages) maps;
        pages[i]: original uint8 HWC CUDA page of map i; scales[i] = (scale_h, scale_w) of resize_image_for_det.
        Queues one kernel on the current stream and returns {"out", "rects", "total"} (device)."""
        if len(pages) != det_buf["counts"].numel() or len(pages) != len(scales):
            raise ValueError("one page and one (scale_h, scale_w) per probability map")
        dev = _need_cuda_page(pages[0]).device
        buf = self._buffers(dev)
        with torch.cuda.device(dev):
            ops.rec_prep_boxes(pages, scales, det_buf["boxes"], det_buf["counts"], max_candidates, buf["out"], buf["rects"],
                               buf["total"], self.mean, self.std)
        buf["pages_ref"] = list(pages)  # keep alive until the stream has consumed them
        return buf

    def from_rects(self, pages: Sequence[torch.Tensor], rects: np.ndarray):
        """rects: host (N, 5) int array [image, x, y, w, h] in page coordinates (slices of the pages)."""
        rects = np.asarray(rects, dtype=np.int32).reshape(-1, 5)
        if len(rects) > self.capacity:
            raise ValueError(f"{len(rects)} crops exceed the capacity {self.capacity}")
        dev = _need_cuda_page(pages[0]).device
        buf = self._buffers(dev)
        rin = np.zeros((max(len(rects), 1), 6), dtype=np.int32)
        rin[:len(rects), :5] = rects
        rin_dev = torch.from_numpy(rin).to(dev, non_blocking=False)
        with torch.cuda.device(dev):
            ops.rec_prep_rects(pages, rin_dev, len(rects), buf["out"], buf["rects"], buf["total"], self.mean, self.std)
        buf["pages_ref"] = list(pages) + [rin_dev]
        return buf

    # ------------------------------------------------------------------ host side
    def fetch(self, buf, on_degenerate: str = "raise"):
        """Synchronises; returns (batch tensor view out[:n] on the device, host structured array of n rectangles).
        `on_degenerate`: "raise" (the reference raises cv2.error) or "zeros" (keep the all-zero tensor)."""
        total = int(buf["total"].item())
        n = min(total, self.capacity)
        rects = buf["rects"][:n * 6].cpu().numpy().view(_RECT_DTYPE).reshape(n)
        if total > self.capacity:
            raise RuntimeError(f"{total} crops exceed the capacity {self.capacity}: raise `capacity`")
        if on_degenerate == "raise" and (rects["status"] == STATUS_DEGENERATE).any():
            bad = int(np.flatnonzero(rects["status"] == STATUS_DEGENERATE)[0])
            raise DegenerateCropError(f"crop {bad} ({rects['w'][bad]} x {rects['h'][bad]}) resizes to zero width")
        return buf["out"][:n], rects


_single = {}


def preprocess_for_recognition(crop: torch.Tensor, img_size: Tuple[int, int] = (32, 256)) -> torch.Tensor:
    """pipeline2.py:92-128 for one uint8 HWC CUDA crop (a view produced by `crop_image` is fine): (3, th, tw) float32
    CUDA tensor.  Raises DegenerateCropError where the reference's cv2.resize raises; an empty crop raises ZeroDivisionError
    (the reference divides by zero at pipeline2.py:101)."""
    _need_cuda_page(crop)
    if crop.numel() == 0:
        raise ZeroDivisionError("preprocess_for_recognition: empty crop (the reference divides by its zero height)")
    if crop.stride(2) != 1 or crop.stride(1) != 3:
        crop = crop.contiguous()
    key = (tuple(img_size), str(crop.device))
    cr = _single.get(key)
    if cr is None:
        cr = _single[key] = RecognitionCropper(img_size, capacity=1, device=crop.device)
    buf = cr.from_rects([crop], np.array([[0, 0, 0, crop.shape[1], crop.shape[0]]], dtype=np.int32))
    out, _ = cr.fetch(buf, on_degenerate="raise")
    return out[0].clone()


def recognition_batch(crops: List[torch.Tensor], img_size: Tuple[int, int] = (32, 256)) -> torch.Tensor:
    """The tensor assembly of recognize_text_batch (pipeline2.py:149-159) for a list of CUDA crops (views of at most 32
    distinct pages are batched into one launch per 32): empty crops give zero tensors, like the reference."""
    th, tw = img_size
    if not crops:
        return torch.empty(0, 3, th, tw)
    dev = crops[0].device
    out = torch.empty(len(crops), 3, th, tw, dtype=torch.float32, device=dev)
    cr = RecognitionCropper(img_size, capacity=_lib.REC_MAX_IMAGES, device=dev)
    for s in range(0, len(crops), _lib.REC_MAX_IMAGES):
        part = crops[s:s + _lib.REC_MAX_IMAGES]
        pages, rects = [], []
        for c in part:
            _need_cuda_page(c)
            if c.numel() == 0:
                rects.append([len(pages) - 1 if pages else 0, 0, 0, 0, 0])
                continue
            if c.stride(2) != 1 or c.stride(1) != 3:
                c = c.contiguous()
            pages.append(c)
            rects.append([len(pages) - 1, 0, 0, c.shape[1], c.shape[0]])
        if not pages:
            out[s:s + len(part)].zero_()
            continue
        got, _ = cr.fetch(cr.from_rects(pages, np.array(rects, dtype=np.int32)), on_degenerate="raise")
        out[s:s + len(part)].copy_(got)
    return out
