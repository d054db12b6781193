"""CPU restatement of the detection -> recognition hand-over -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

SURVEY.md 8(f) item 2.  Reference code restated here (file:line under the reference tree):
  * box rescale to the original image           src/pipeline/pipeline2.py:323-328
  * `crop_image`                                src/det/test.py:123-130   (cv2.boundingRect + clamps + slice)
  * `preprocess_for_recognition`                src/pipeline/pipeline2.py:92-128
  * the blank tensor for an empty crop          src/pipeline/pipeline2.py:153-157

Third-party arithmetic: `cv2.boundingRect` of an integer point set (min / max + 1) and `cv2.resize` (8-bit
INTER_LINEAR, restated in oracle/preprocess_oracle.py and pinned there); `cv2.copyMakeBorder` with a constant is a
plain fill.  OpenCV is unpinned by the reference (4.13.0 installed in the authoring container).

PINNED: tests/golden/recprep_reference.npz holds outputs of the reference's own `crop_image` and
`preprocess_for_recognition` (imported unchanged from /root/reference by tests/golden/make_recprep_golden.py, cv2 and
torch executing as written) for boxes that are interior, clipped by every image border, negative, fully outside,
degenerate, wider than the 256-pixel target and narrower than one output pixel; this restatement matches them bit for
bit (tests/test_oracle.py).

Reference quirks kept on purpose:
  * `box[:, 0] = box[:, 0] / scale_w` assigns float64 quotients into an int64 array: truncation toward zero, not floor;
  * `crop_image` clamps a negative x to 0 WITHOUT shrinking the width (the crop is shifted, not cut);
  * `new_w = int(w * (target_h / h))` with the quotient rounded first; `new_w == 0` makes `cv2.resize` raise
    (`inv_scale_x > 0`): status DEGENERATE here, the caller decides;
  * an empty crop becomes an all-ZERO tensor, not a normalised white one (pipeline2.py:155-157).
"""
from __future__ import annotations

import numpy as np

from .preprocess_oracle import resize_u8_linear

STATUS_OK = 0
STATUS_EMPTY = 1       # crop.size == 0 -> torch.zeros(3, th, tw)        (pipeline2.py:155-157)
STATUS_DEGENERATE = 2  # new_w == 0 -> cv2.resize raises in the reference (pipeline2.py:106)

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def rescale_box(box: np.ndarray, scale_h: float, scale_w: float) -> np.ndarray:
    """pipeline2.py:323-328: in-place division of an integer (K, 2) array by python floats, then astype(int32)."""
    box = np.array(box, dtype=np.int64)
    box[:, 0] = np.trunc(box[:, 0].astype(np.float64) / np.float64(scale_w)).astype(np.int64)
    box[:, 1] = np.trunc(box[:, 1].astype(np.float64) / np.float64(scale_h)).astype(np.int64)
    return box.astype(np.int32)


def bounding_rect(box: np.ndarray):
    """cv2.boundingRect for an integer point set: (min x, min y, max x - min x + 1, max y - min y + 1)."""
    x0, y0 = int(box[:, 0].min()), int(box[:, 1].min())
    x1, y1 = int(box[:, 0].max()), int(box[:, 1].max())
    return x0, y0, x1 - x0 + 1, y1 - y0 + 1


def crop_rect(img_h: int, img_w: int, box: np.ndarray):
    """The (x, y, bw, bh) `crop_image` slices with (test.py:124-130); bw / bh may be <= 0 (empty crop)."""
    x, y, bw, bh = bounding_rect(box)
    x = max(0, x)
    y = max(0, y)
    bw = min(bw, img_w - x)
    bh = min(bh, img_h - y)
    return x, y, bw, bh


def crop_image(img: np.ndarray, box: np.ndarray) -> np.ndarray:
    x, y, bw, bh = crop_rect(img.shape[0], img.shape[1], box)
    if bw <= 0 or bh <= 0:
        return img[0:0, 0:0]
    return img[y:y + bh, x:x + bw]


def rec_width(h: int, w: int, target_h: int = 32) -> int:
    """pipeline2.py:101-102 (float64 quotient rounded before the product, int() truncates)."""
    return int(np.float64(w) * (np.float64(target_h) / np.float64(h)))


def preprocess_for_recognition(crop: np.ndarray, img_size=(32, 256), mean=MEAN, std=STD):
    """Returns (float32 (3, th, tw), status).  crop: uint8 (h, w, 3)."""
    th, tw = img_size
    if crop.size == 0:
        return np.zeros((3, th, tw), dtype=np.float32), STATUS_EMPTY
    h, w = crop.shape[:2]
    new_w = rec_width(h, w, th)
    if new_w == 0:
        return np.zeros((3, th, tw), dtype=np.float32), STATUS_DEGENERATE
    if new_w > tw:
        resized = resize_u8_linear(np.ascontiguousarray(crop), tw, th)
    else:
        resized = np.full((th, tw, 3), 255, dtype=np.uint8)
        resized[:, :new_w] = resize_u8_linear(np.ascontiguousarray(crop), new_w, th)
    x = resized.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)
    x = (x - np.asarray(mean, dtype=np.float32).reshape(3, 1, 1)) / np.asarray(std, dtype=np.float32).reshape(3, 1, 1)
    return x.astype(np.float32), STATUS_OK


def crops_for_recognition(img: np.ndarray, boxes, scale_h: float, scale_w: float, img_size=(32, 256)):
    """Detector boxes (resized-image coordinates) -> (N, 3, th, tw) float32, (N, 5) int32 [x, y, bw, bh, status]."""
    th, tw = img_size
    out = np.zeros((len(boxes), 3, th, tw), dtype=np.float32)
    rects = np.zeros((len(boxes), 5), dtype=np.int32)
    for i, box in enumerate(boxes):
        rb = rescale_box(box, scale_h, scale_w)
        x, y, bw, bh = crop_rect(img.shape[0], img.shape[1], rb)
        out[i], st = preprocess_for_recognition(crop_image(img, rb), img_size)
        rects[i] = (x, y, bw, bh, st)
    return out, rects
