"""Golden vectors for the post-processing from the REFERENCE'S OWN CODE (run in the authoring container).

    python tests/golden/make_postproc_reference.py

Imports /root/reference/src/det/test.py unchanged and runs its `DBPostProcessor.__call__` (test.py:55-106, with
`box_score_fast` :20-34 and `unclip` :37-43) on rendered probability maps.  The module-level imports of that file
that are absent offline are satisfied with stand-ins registered in sys.modules BEFORE the import:

  * `pyclipper`, `shapely.geometry`  -> oracle/clipper_offset.py, a restatement of Clipper 6.4.2's ClipperOffset
    (the final self-intersection union is not restated: cv2.boundingRect of a box -- the only thing the reference's
    consumer reads, test.py:123-130 -- is unaffected);
  * `matplotlib.pyplot`, `src.preprocess.scanner` (imports rembg)  -> empty stubs: neither is touched by the
    post-processor.

Everything else (cv2.findContours / approxPolyDP / fillPoly / mean / contourArea, the loop, the filters, the order
of the results) is the reference's code executing as written.  Stored per map: the probability map, and for every
returned box its cv2.boundingRect (x, y, w, h), vertex count and score, in the reference's output order.
/root/reference does not exist on the GPU box, so the tests only read the .npz.
"""
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from oracle import clipper_offset  # noqa: E402
from oracle import postproc_oracle as ppo  # noqa: E402


def _install_stand_ins():
    pyclipper = types.ModuleType("pyclipper")
    for name in ("PyclipperOffset", "JT_ROUND", "JT_SQUARE", "JT_MITER", "ET_CLOSEDPOLYGON"):
        setattr(pyclipper, name, getattr(clipper_offset, name))
    sys.modules["pyclipper"] = pyclipper
    shapely = types.ModuleType("shapely")
    geometry = types.ModuleType("shapely.geometry")
    geometry.Polygon = clipper_offset.Polygon
    shapely.geometry = geometry
    sys.modules["shapely"] = shapely
    sys.modules["shapely.geometry"] = geometry
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    scanner = types.ModuleType("src.preprocess.scanner")
    scanner.preprocess_image = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("scanner stub"))
    sys.modules["src.preprocess.scanner"] = scanner


def clean_rect_map(h, w, seed):
    """Axis-aligned constant-probability rectangles, no noise: contour polygon == pixel-centre rectangle."""
    rng = np.random.default_rng(seed)
    prob = np.full((h, w), 0.05, dtype=np.float32)
    occ = np.zeros((h, w), dtype=bool)
    placed = 0
    for _ in range(400):
        bh, bw = int(rng.integers(5, 24)), int(rng.integers(8, 90))
        y, x = int(rng.integers(1, h - bh - 1)), int(rng.integers(1, w - bw - 1))
        if occ[max(0, y - 4):y + bh + 4, max(0, x - 4):x + bw + 4].any():
            continue
        level = [0.95, 0.8, 0.7, 0.55][placed % 4]
        prob[y:y + bh, x:x + bw] = level
        occ[y:y + bh, x:x + bw] = True
        placed += 1
        if placed == 40:
            break
    return prob


def main():
    import cv2
    _install_stand_ins()
    from src.det.test import DBPostProcessor  # the reference, unchanged

    maps, params = [], []
    for seed in (11, 12):
        maps.append(clean_rect_map(256, 384, seed)); params.append((0.3, 0.6, 1.5, 10, 1))
    for seed, ratio, bt in ((3, 1.5, 0.6), (4, 1.6, 0.5), (5, 2.0, 0.6)):
        maps.append(ppo.render_prob_map(256, 384, seed=seed, n_boxes=28)); params.append((0.3, bt, ratio, 10, 0))
    out = {"n_maps": np.int32(len(maps))}
    for i, (prob, (th, bt, ratio, min_area, clean)) in enumerate(zip(maps, params)):
        post = DBPostProcessor(thresh=th, box_thresh=bt, max_candidates=1000, unclip_ratio=ratio)
        post.min_area = min_area
        boxes, scores = post(prob[None])
        rects = np.array([cv2.boundingRect(np.asarray(b).astype(np.int32)) for b in boxes], dtype=np.int32).reshape(-1, 4)
        out[f"prob_{i}"] = prob
        out[f"params_{i}"] = np.array([th, bt, ratio, min_area, clean], dtype=np.float64)
        out[f"rects_{i}"] = rects
        out[f"nverts_{i}"] = np.array([len(b) for b in boxes], dtype=np.int32)
        out[f"scores_{i}"] = np.array(scores, dtype=np.float64)
        print(f"map {i}: {len(boxes)} boxes from the reference post-processor, vertices {[len(b) for b in boxes][:6]}...")
    np.savez_compressed(HERE / "postproc_reference.npz", **out)
    print("wrote", HERE / "postproc_reference.npz")


if __name__ == "__main__":
    main()
