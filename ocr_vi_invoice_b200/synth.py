"""Synthetic invoice-like inputs for benchmarks and smoke runs.

The reference's generator (generator/det, ~3.9 k lines of PIL layouts) is an offline producer of JPEG + JSON
and is not rebuilt; it is also unavailable on the GPU box.  This is a seeded numpy
stand-in with the properties the detection path cares about: a light page, dark axis-aligned text-line
rectangles made of character-like strokes, the shape the caller asks for, uint8 RGB -- then the reference's own
normalisation (src/pipeline/pipeline2.py:312-314: /255, ImageNet mean/std, HWC -> CHW float32).
"""
from __future__ import annotations

import numpy as np

MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def synthetic_invoice(h: int, w: int, seed: int = 0, n_lines: int = 60):
    """Returns (uint8 RGB image (h, w, 3), [(x0, y0, x1, y1)] text boxes)."""
    rng = np.random.default_rng(seed)
    img = np.full((h, w, 3), 255, dtype=np.uint8)
    img[:] = rng.integers(235, 256, size=3, dtype=np.uint8)
    img = np.clip(img.astype(np.int16) + rng.integers(-6, 7, size=(h, w, 1)), 0, 255).astype(np.uint8)
    boxes = []
    y = int(0.03 * h)
    while y < h - 30 and len(boxes) < n_lines:
        lh = int(rng.integers(max(10, h // 70), max(12, h // 28)))
        x = int(rng.integers(int(0.03 * w), int(0.15 * w)))
        while x < w - 60:
            lw = int(rng.integers(w // 16, w // 3))
            x1 = min(w - 8, x + lw)
            ink = int(rng.integers(10, 90))
            cx = x
            while cx < x1 - 3:  # character-like strokes
                cw = int(rng.integers(max(3, lh // 3), max(4, lh)))
                if rng.random() > 0.15:
                    img[y + int(rng.integers(0, 3)):y + lh - int(rng.integers(0, 3)), cx:min(x1, cx + cw - 2)] = ink
                cx += cw
            boxes.append((x, y, x1, y + lh))
            x = x1 + int(rng.integers(w // 30, w // 6))
        y += lh + int(rng.integers(lh // 2, 2 * lh))
    return img, boxes


def normalize(img_u8: np.ndarray) -> np.ndarray:
    """uint8 HWC RGB -> float32 CHW, the reference's preprocessing (pipeline2.py:312-314)."""
    x = img_u8.astype(np.float32) / 255.0
    x = (x - MEAN) / STD
    return np.ascontiguousarray(x.transpose(2, 0, 1), dtype=np.float32)


def synthetic_batch(batch: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """(batch, 3, h, w) float32 normalised batch of distinct synthetic invoices."""
    return np.stack([normalize(synthetic_invoice(h, w, seed * 1000 + i)[0]) for i in range(batch)])


def synthetic_batch_u8(batch: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """(batch, h, w, 3) uint8 RGB batch of the same synthetic invoices (the decoded-image interface)."""
    return np.stack([synthetic_invoice(h, w, seed * 1000 + i)[0] for i in range(batch)])
