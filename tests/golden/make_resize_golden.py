"""Golden vectors for the image resize: outputs of cv2.resize (the call of resize_image_for_det,
/root/reference/src/pipeline/pipeline2.py:39) on seeded uint8 images, run in the authoring container (cv2 4.13.0).

    python tests/golden/make_resize_golden.py
"""
from pathlib import Path

import cv2
import numpy as np

HERE = Path(__file__).resolve().parent
CASES = [(50, 70, 224, 96), (120, 90, 64, 96), (97, 131, 160, 128), (7, 9, 64, 32), (64, 64, 32, 32), (3, 3, 96, 96),
         (200, 150, 160, 128), (1, 5, 32, 64)]


def main():
    rng = np.random.default_rng(2024)
    out = {"cases": np.array(CASES, dtype=np.int32), "cv2_version": np.array(cv2.__version__)}
    for i, (h, w, dh, dw) in enumerate(CASES):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if i % 2:  # smooth content as well as noise
            img = cv2.GaussianBlur(img, (5, 5), 0)
        out[f"src_{i}"] = img
        out[f"dst_{i}"] = cv2.resize(img, (dw, dh))
    np.savez_compressed(HERE / "resize_cv2.npz", **out)
    print("wrote", HERE / "resize_cv2.npz")


if __name__ == "__main__":
    main()
