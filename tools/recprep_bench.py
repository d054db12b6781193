"""Microbench of the detection -> recognition hand-over kernel (b200_rec_prep_boxes) on one B200.

    python tools/recprep_bench.py [--pages 16] [--boxes 60] [--iters 50]

Workload: `pages` synthetic 1400 x 1000 uint8 pages (the generator's largest native size), `boxes`
text-line boxes per page in detector coordinates (1024-pixel long side), one launch for the whole batch.
Algorithmic bytes per launch = the source bytes of every crop (read once) + 3 * 32 * 256 * 4 output bytes per crop;
reported against the measured HBM copy peak of MEASURED_PEAKS.json.  Timed with CUDA events on the launching stream,
L2 flushed between iterations.
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=16)
    ap.add_argument("--boxes", type=int, default=60)
    ap.add_argument("--iters", type=int, default=50)
    args = ap.parse_args()
    from ocr_vi_invoice_b200.postprocess import _BOX_DTYPE
    from ocr_vi_invoice_b200.preprocess import det_size
    from ocr_vi_invoice_b200.recprep import RecognitionCropper

    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    h, w = 1400, 1000
    (nh, nw), (sh, sw) = det_size(h, w, 1024)
    pages = [torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device=dev) for _ in range(args.pages)]
    mc = 1000
    recs = np.zeros((args.pages, mc), dtype=_BOX_DTYPE)
    src_bytes = 0
    for i in range(args.pages):
        for j in range(args.boxes):
            bw, bh = int(rng.integers(60, 500)), int(rng.integers(18, 48))
            x0, y0 = int(rng.integers(0, nw - bw)), int(rng.integers(0, nh - bh))
            recs[i, j]["x0"], recs[i, j]["y0"], recs[i, j]["x1"], recs[i, j]["y1"] = x0, y0, x0 + bw, y0 + bh
            src_bytes += int((bw / sw + 1)) * int((bh / sh + 1)) * 3
    det = {"boxes": torch.from_numpy(recs.view(np.uint8).reshape(-1)).to(dev),
           "counts": torch.full((args.pages,), args.boxes, dtype=torch.int32, device=dev)}
    n = args.pages * args.boxes
    cr = RecognitionCropper((32, 256), capacity=max(1024, n))
    scales = [(sh, sw)] * args.pages
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        cr.from_detector(det, pages, scales, mc)
    torch.cuda.synchronize()
    ms = []
    for _ in range(args.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cr.from_detector(det, pages, scales, mc)
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms))
    out_bytes = n * 3 * 32 * 256 * 4
    peak = 6448.4
    try:
        peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
    except Exception:
        pass
    gbs = (src_bytes + out_bytes) / t / 1e6
    print(json.dumps({"kernel": "rec_prep_kernel<true>", "pages": args.pages, "crops": n, "ms": round(t, 4),
                      "crops_per_s": round(n / t * 1e3), "algorithmic_bytes": src_bytes + out_bytes,
                      "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3)}))


if __name__ == "__main__":
    main()
