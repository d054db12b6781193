"""CPU restatement of the detection post-processing contract -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference post-processor (DBPostProcessor.__call__, src/det/test.py:55-106) walks cv2 contours in a
Python loop and un-clips each polygon with pyclipper.  BASELINE.json's north star replaces it by threshold ->
connected-component labelling -> per-component box extraction and mean-score filtering on the GPU, with the
label map and boxes bit-exact "given identical binarised maps".  This file is the CPU statement of exactly
that contract; the CUDA path in ocr_vi_invoice_b200/csrc/postproc.cu must reproduce
it bit for bit.

Pinning.  The reference's own `DBPostProcessor.__call__` was run in the authoring container on rendered maps
(tests/golden/make_postproc_reference.py: src/det/test.py imported unchanged; pyclipper / shapely, which are not
installable here, replaced by oracle/clipper_offset.py, a restatement of Clipper 6.4.2's ClipperOffset without its
final self-intersection union -- the bounding rectangle of a box, all the reference's consumer reads, is unaffected).
Against those goldens (tests/golden/postproc_reference.npz, tests/test_oracle.py):
  * `cv2_reference_path` (the restated cv2 half + closed-form bounding-rectangle growth) reproduces the reference's
    bounding rectangles and scores exactly, in order;
  * this contract's boxes and scores EQUAL the reference's on maps whose components are axis-aligned rectangles;
  * on noisy maps with holes / diagonal touches they agree for > 90 % of the boxes (IoU >= 0.8) -- the documented
    differences: hole contours (RETR_LIST), polygon-mask mean vs component-pixel mean, approxPolyDP corners.
PARITY PARTIALLY PINNED: Clipper's union step and the many-vertex polygon outputs are not reproduced.

Contract
  1. foreground = prob > float32(thresh)                                   (test.py:57)
  2. 8-connected components (cv2.findContours traces 8-connected foreground, test.py:60); canonical label =
     1-based rank of the component's first pixel in raster order, background 0
  3. only the first `max_candidates` components (label order) are examined (test.py:66-67)
  4. per component: inclusive bounding box, pixel count, qsum = sum rint(prob * 2^24) (exact integer);
     score = qsum / (npix * 2^24)  replaces box_score_fast's polygon-mask mean (test.py:20-34)
  5. keep iff float64(qsum) >= (box_thresh * 2^24) * npix  (test.py:77)  and the pixel-centre rectangle
     (x0,y0)-(x1,y1) has width >= 1, height >= 1 and area (x1-x0)*(y1-y0) >= min_area (test.py:73,81; a cv2
     contour runs through pixel centres, so a w x h pixel rectangle has area (w-1)(h-1))
  6. un-clip (test.py:37-43) in closed form for the rectangle: d = area * ratio / perimeter, corners moved
     outward by d and rounded half away from zero (Clipper's Round()); output 4-point polygon
     [(x0',y0'), (x1',y0'), (x1',y1'), (x0',y1')], int64, in label order.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

Q = 16777216.0  # 2^24


def foreground(prob: np.ndarray, thresh: float) -> np.ndarray:
    return np.asarray(prob, dtype=np.float32) > np.float32(thresh)


def label_components(fg: np.ndarray) -> Tuple[np.ndarray, int]:
    """Canonical 8-connected labels via scipy.ndimage (C speed), re-ranked by first pixel in raster order."""
    from scipy import ndimage
    lab, n = ndimage.label(fg, structure=np.ones((3, 3), dtype=np.int32))
    if n == 0:
        return np.zeros(fg.shape, dtype=np.int32), 0
    flat = lab.ravel()
    idx = np.nonzero(flat)[0]
    first = np.full(n + 1, flat.size, dtype=np.int64)
    np.minimum.at(first, flat[idx], idx)
    order = np.argsort(first[1:], kind="stable")  # component ids (0-based) by first pixel
    rank = np.zeros(n + 1, dtype=np.int32)
    rank[order + 1] = np.arange(1, n + 1, dtype=np.int32)
    return rank[lab].astype(np.int32), n


def label_components_numpy(fg: np.ndarray) -> Tuple[np.ndarray, int]:
    """Pure-numpy min-label propagation (slow; cross-checks the scipy path on small maps)."""
    h, w = fg.shape
    big = h * w + 1
    lab = np.where(fg, np.arange(1, h * w + 1, dtype=np.int64).reshape(h, w), big)
    while True:
        p = np.pad(lab, 1, constant_values=big)
        m = lab.copy()
        for dy in (0, 1, 2):
            for dx in (0, 1, 2):
                m = np.minimum(m, p[dy:dy + h, dx:dx + w])
        m = np.where(fg, m, big)
        if np.array_equal(m, lab):
            break
        lab = m
    roots = np.unique(lab[fg])
    out = np.zeros((h, w), dtype=np.int32)
    if roots.size:
        out[fg] = (np.searchsorted(roots, lab[fg]) + 1).astype(np.int32)
    return out, int(roots.size)


def round_away(v: np.ndarray) -> np.ndarray:
    v = np.asarray(v, dtype=np.float64)
    return np.trunc(np.where(v < 0, v - 0.5, v + 0.5)).astype(np.int64)


def postprocess(prob: np.ndarray, thresh: float = 0.3, box_thresh: float = 0.6, max_candidates: int = 1000,
                unclip_ratio: float = 1.5, min_area: int = 10):
    """prob: (H, W) float32.  Returns (labels int32 (H, W), ncomp, records) where records is a structured array
    with fields x0,y0,x1,y1 (un-clipped rectangle), bx0,by0,bx1,by1, label, npix, qsum in label order."""
    prob = np.ascontiguousarray(prob, dtype=np.float32)
    fg = foreground(prob, thresh)
    labels, n = label_components(fg)
    dt = np.dtype([("x0", "<i4"), ("y0", "<i4"), ("x1", "<i4"), ("y1", "<i4"), ("bx0", "<i4"), ("by0", "<i4"),
                   ("bx1", "<i4"), ("by1", "<i4"), ("label", "<i4"), ("npix", "<i4"), ("qsum", "<u8")])
    nc = min(n, max_candidates)
    if nc == 0:
        return labels, n, np.zeros(0, dtype=dt)
    ys, xs = np.nonzero(labels)
    lb = labels[ys, xs].astype(np.int64)
    sel = lb <= nc
    ys, xs, lb = ys[sel], xs[sel], lb[sel]
    q = np.rint(prob[ys, xs] * np.float32(Q)).astype(np.uint64)  # exact: power-of-two scaling in fp32
    npix = np.bincount(lb, minlength=nc + 1).astype(np.int64)
    qsum = np.zeros(nc + 1, dtype=np.uint64)
    np.add.at(qsum, lb, q)
    minx = np.full(nc + 1, np.iinfo(np.int64).max); maxx = np.full(nc + 1, -1)
    miny = np.full(nc + 1, np.iinfo(np.int64).max); maxy = np.full(nc + 1, -1)
    np.minimum.at(minx, lb, xs); np.maximum.at(maxx, lb, xs)
    np.minimum.at(miny, lb, ys); np.maximum.at(maxy, lb, ys)
    recs = []
    tq = np.float64(box_thresh) * Q
    for c in range(1, nc + 1):
        bw, bh = int(maxx[c] - minx[c]), int(maxy[c] - miny[c])
        area = bw * bh
        score_ok = np.float64(qsum[c]) >= tq * np.float64(npix[c])
        if not (score_ok and bw >= 1 and bh >= 1 and area >= min_area):
            continue
        d = (np.float64(area) * np.float64(unclip_ratio)) / np.float64(2 * (bw + bh))
        x0, y0 = round_away(minx[c] - d), round_away(miny[c] - d)
        x1, y1 = round_away(maxx[c] + d), round_away(maxy[c] + d)
        recs.append((x0, y0, x1, y1, minx[c], miny[c], maxx[c], maxy[c], c, npix[c], qsum[c]))
    return labels, n, np.array(recs, dtype=dt)


def records_to_boxes(recs) -> Tuple[List[np.ndarray], List[float]]:
    """Box list in the shape DBPostProcessor returns (test.py:106): list[(4,2) int64], list[float]."""
    boxes = [np.array([[r["x0"], r["y0"]], [r["x1"], r["y0"]], [r["x1"], r["y1"]], [r["x0"], r["y1"]]], dtype=np.int64)
             for r in recs]
    scores = [float(np.float64(r["qsum"]) / (np.float64(r["npix"]) * Q)) for r in recs]
    return boxes, scores


# ----------------------------------------------------------------------------- fixtures / reported agreement
def render_prob_map(h: int, w: int, seed: int = 0, n_boxes: int = 32, noise: float = 0.05) -> np.ndarray:
    """Synthetic probability map: text-line-like rectangles (some touching edge-to-edge, some only diagonally,
    one with a hole, a few low-score ones) over a noisy low background."""
    rng = np.random.default_rng(seed)
    prob = np.clip(0.05 + noise * rng.standard_normal((h, w)), 0.0, 0.25).astype(np.float32)
    placed = 0
    tries = 0
    occupied = np.zeros((h, w), dtype=bool)
    while placed < n_boxes and tries < 50 * n_boxes:
        tries += 1
        bh = int(rng.integers(6, max(7, h // 12)))
        bw = int(rng.integers(12, max(13, w // 3)))
        y = int(rng.integers(0, max(1, h - bh)))
        x = int(rng.integers(0, max(1, w - bw)))
        y2, x2 = min(h, y + bh + 3), min(w, x + bw + 3)
        if occupied[max(0, y - 3):y2, max(0, x - 3):x2].any():
            continue
        level = 0.9 if placed % 7 else 0.45  # every 7th box is faint: passes thresh, fails box_thresh
        patch = np.clip(level + noise * rng.standard_normal((bh, bw)), 0.0, 1.0)
        prob[y:y + bh, x:x + bw] = patch
        occupied[y:y + bh, x:x + bw] = True
        placed += 1
    # structured edge cases in the top-left corner region (if room)
    if h >= 64 and w >= 96:
        prob[2:10, 2:20] = 0.85; prob[10:18, 20:40] = 0.85        # diagonal touch -> one 8-connected component
        prob[24:40, 4:44] = 0.9; prob[28:36, 12:36] = 0.02        # ring with a hole
        prob[44, 4:30] = 0.95                                     # single-row run: zero-area, dropped
        prob[48:52, 60] = 0.95                                    # single column
        prob[56, 70] = 0.99                                       # isolated pixel
    return prob.astype(np.float32)


def cv2_reference_path(prob: np.ndarray, thresh=0.3, box_thresh=0.6, max_candidates=1000, unclip_ratio=1.5, min_area=10):
    """The cv2 half of DBPostProcessor.__call__ (src/det/test.py:55-106) restated, with `unclip`
    (pyclipper JT_ROUND offset, test.py:37-43) replaced by its effect on the polygon's bounding rectangle: the
    consumer only takes cv2.boundingRect(box) (test.py:123-130), and a round-joined offset by d grows the bounding
    rectangle by exactly d on every side.  Returns ([(x0, y0, x1, y1)], [score]) in the reference's contour order."""
    import cv2
    seg = np.asarray(prob, dtype=np.float32) > np.float32(thresh)
    contours, _ = cv2.findContours((seg * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    rects, scores = [], []
    hh, ww = prob.shape
    for i, contour in enumerate(contours):
        if i >= max_candidates:
            break
        eps = 0.002 * cv2.arcLength(contour, True)
        pts = cv2.approxPolyDP(contour, eps, True).reshape(-1, 2)
        if pts.shape[0] < 4:
            continue
        xmin = int(np.clip(pts[:, 0].min(), 0, ww - 1)); xmax = int(np.clip(pts[:, 0].max(), 0, ww - 1))
        ymin = int(np.clip(pts[:, 1].min(), 0, hh - 1)); ymax = int(np.clip(pts[:, 1].max(), 0, hh - 1))
        mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
        cv2.fillPoly(mask, [(pts - np.array([xmin, ymin])).astype(np.int32)], 1)
        score = cv2.mean(prob[ymin:ymax + 1, xmin:xmax + 1], mask)[0]
        if box_thresh > score:
            continue
        area = cv2.contourArea(pts)
        if area < min_area:
            continue
        d = area * unclip_ratio / cv2.arcLength(pts.reshape(-1, 1, 2), True)
        rects.append((int(round_away(xmin - d)), int(round_away(ymin - d)), int(round_away(xmax + d)),
                      int(round_away(ymax + d))))
        scores.append(float(score))
    return rects, scores


def rect_iou(a, b) -> float:
    ix = max(0, min(a[2], b[2]) - max(a[0], b[0]) + 1)
    iy = max(0, min(a[3], b[3]) - max(a[1], b[1]) + 1)
    inter = ix * iy
    ua = (a[2] - a[0] + 1) * (a[3] - a[1] + 1) + (b[2] - b[0] + 1) * (b[3] - b[1] + 1) - inter
    return inter / ua if ua > 0 else 0.0


def match_rate(ours, theirs, iou: float = 0.8) -> float:
    """Fraction of `theirs` rectangles that have a counterpart in `ours` with IoU >= iou."""
    if not theirs:
        return 1.0
    hit = sum(1 for t in theirs if any(rect_iou(t, o) >= iou for o in ours))
    return hit / len(theirs)
