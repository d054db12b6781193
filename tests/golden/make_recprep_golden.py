"""Golden vectors for the detection -> recognition hand-over from the REFERENCE'S OWN CODE (authoring container only).

    python tests/golden/make_recprep_golden.py

Imports /root/reference/src/pipeline/pipeline2.py and src/det/test.py unchanged (the modules absent offline are
stubbed exactly as in make_postproc_reference.py; none is touched by these functions) and runs, as written:
  * the box rescale loop of pipeline2.py:323-328,
  * `crop_image` (test.py:123-130, cv2.boundingRect),
  * `preprocess_for_recognition` (pipeline2.py:92-128, cv2.resize / copyMakeBorder / torch normalisation),
  * the blank-tensor branch of `recognize_text_batch` (pipeline2.py:153-157),
on a seeded synthetic RGB page and a list of detector boxes covering: interior boxes of many aspect ratios, boxes cut
by each image border, negative corners, boxes fully outside, one-pixel boxes, crops wider than 256 output pixels and
crops so tall that new_w == 0 (the reference raises cv2.error there: recorded as status 2).

Stored: the page, the boxes, the scales, and per box the crop rectangle, the status and the returned float32
(3, 32, 256) tensor -- the latter as byte indices into `norm_table`, the reference's own output for every (byte value,
channel) pair: the generator asserts that table[c][index] reproduces the returned tensor bit for bit, so the fixture
stays small without losing a bit.  /root/reference does not exist on the GPU box, so the tests only read the .npz.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))
sys.path.insert(0, "/root/reference")

import make_postproc_reference as stand_ins  # noqa: E402


def synthetic_page(h, w, seed):
    """Text-like page: light background, dark glyph-sized blobs, per-pixel noise, a colour gradient."""
    rng = np.random.default_rng(seed)
    img = np.full((h, w, 3), 235, dtype=np.float32)
    img += np.linspace(0, -40, w, dtype=np.float32)[None, :, None] * np.array([1.0, 0.5, 0.2], dtype=np.float32)
    for _ in range(h * w // 800):
        gh, gw = int(rng.integers(3, 14)), int(rng.integers(2, 10))
        y, x = int(rng.integers(0, h - gh)), int(rng.integers(0, w - gw))
        img[y:y + gh, x:x + gw] = rng.integers(0, 120, size=3)
    img += rng.normal(0, 6, size=img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def rect_box(x0, y0, x1, y1):
    return np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], dtype=np.int64)


def box_list(rh, rw, seed):
    """Boxes in RESIZED-image coordinates (what DBPostProcessor returns), rh x rw = detector input size."""
    rng = np.random.default_rng(seed)
    boxes = []
    for _ in range(14):  # interior, many aspect ratios
        bw, bh = int(rng.integers(6, 420)), int(rng.integers(5, 60))
        x0, y0 = int(rng.integers(0, rw - bw)), int(rng.integers(0, rh - bh))
        boxes.append(rect_box(x0, y0, x0 + bw, y0 + bh))
    boxes += [
        rect_box(-9, 40, 120, 70),            # negative x: shifted, not cut (reference quirk)
        rect_box(30, -7, 200, 25),            # negative y
        rect_box(-20, -20, 60, 15),           # both
        rect_box(rw - 50, 100, rw + 30, 140),  # cut by the right border
        rect_box(100, rh - 20, 300, rh + 40),  # cut by the bottom border
        rect_box(rw + 5, 10, rw + 80, 40),    # fully outside (right)  -> empty
        rect_box(10, rh + 3, 80, rh + 30),    # fully outside (below)  -> empty
        rect_box(-80, 10, -5, 40),            # fully outside (left): x clamps to 0, the width stays -> NOT empty
        rect_box(50, 50, 50, 50),             # one pixel
        rect_box(70, 60, 71, 400),            # 2 px wide, very tall: new_w == 0 -> cv2.resize raises
        rect_box(0, 0, rw - 1, 18),           # full width, thin: new_w > 256 -> both axes scaled
        rect_box(200, 300, 203, 340),         # tall and thin, new_w >= 1
        rect_box(5, 5, 36, 36),               # exactly square
        rect_box(300, 200, 555, 231),         # 256 x 32 in the resized image
    ]
    # a general quadrilateral (is_output_polygon-style input): only its bounding rectangle matters
    boxes.append(np.array([[400, 400], [520, 395], [530, 440], [405, 452]], dtype=np.int64))
    return boxes


def main():
    import cv2
    import torch
    stand_ins._install_stand_ins()
    import src.pipeline.pipeline2 as p2  # the reference, unchanged
    from src.det.test import crop_image

    out = {}
    # the normalisation table: every (byte value, channel) through the reference's torch arithmetic
    ramp = np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 3, axis=2)  # (1, 256, 3)
    ramp = np.repeat(ramp, 32, axis=0)                                            # (32, 256, 3): identity resize
    table = p2.preprocess_for_recognition(ramp, (32, 256)).numpy()[:, 0, :]      # (3, 256)
    assert all(np.all(np.diff(table[c]) > 0) for c in range(3))
    out["norm_table"] = table

    def to_index(t):
        idx = np.stack([np.searchsorted(table[c], t[c]) for c in range(3)]).astype(np.uint8)
        assert np.array_equal(np.stack([table[c][idx[c]] for c in range(3)]), t)
        return idx

    cases = [("a", 620, 438, 640, 11), ("b", 260, 760, 1024, 12), ("c", 167, 167, 640, 13)]
    for name, h, w, det_size, seed in cases:
        page = synthetic_page(h, w, seed)
        resized, (scale_h, scale_w) = p2.resize_image_for_det(page, det_size)
        boxes = box_list(resized.shape[0], resized.shape[1], seed + 100)
        rects, status, tensors = [], [], []
        for box in boxes:
            box = box.copy()
            # pipeline2.py:323-328, verbatim
            box[:, 0] = box[:, 0] / scale_w
            box[:, 1] = box[:, 1] / scale_h
            rb = box.astype(np.int32)
            crop = crop_image(page, rb)
            x, y, bw, bh = cv2.boundingRect(rb)
            x, y = max(0, x), max(0, y)
            rects.append((x, y, min(bw, w - x), min(bh, h - y)))
            if crop.size == 0:
                assert not torch.zeros(3, 32, 256).numpy().any()  # pipeline2.py:155-157
                tensors.append(np.zeros((3, 32, 256), dtype=np.uint8))
                status.append(1)
                continue
            try:
                t = p2.preprocess_for_recognition(crop, (32, 256)).numpy()
                assert t.shape == (3, 32, 256) and t.dtype == np.float32
                tensors.append(to_index(t))
                status.append(0)
            except cv2.error:
                tensors.append(np.zeros((3, 32, 256), dtype=np.uint8))
                status.append(2)
        out[f"{name}_page"] = page
        out[f"{name}_boxes"] = np.stack([b if b.shape == (4, 2) else b for b in boxes])
        out[f"{name}_scales"] = np.array([scale_h, scale_w], dtype=np.float64)
        out[f"{name}_rects"] = np.array(rects, dtype=np.int32)
        out[f"{name}_status"] = np.array(status, dtype=np.int32)
        out[f"{name}_tensor_index"] = np.stack(tensors)
        print(name, page.shape, "scales", scale_h, scale_w, "status counts", np.bincount(status, minlength=3))
    np.savez_compressed(HERE / "recprep_reference.npz", **out)
    print("wrote", HERE / "recprep_reference.npz")


if __name__ == "__main__":
    main()
