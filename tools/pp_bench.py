"""Times the GPU post-processing on batch-16 1024x1024 maps of different character (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ocr_vi_invoice_b200.postprocess import DBPostProcessor


def render_prob_map(h, w, seed=0, n_boxes=150, noise=0.05):
    """Text-line-like rectangles over a noisy low background (an input generator only; the checker lives in oracle/)."""
    rng = np.random.default_rng(seed)
    prob = np.clip(0.05 + noise * rng.standard_normal((h, w)), 0.0, 0.25).astype(np.float32)
    occ = np.zeros((h, w), dtype=bool)
    placed = tries = 0
    while placed < n_boxes and tries < 50 * n_boxes:
        tries += 1
        bh, bw = int(rng.integers(6, h // 12)), int(rng.integers(12, w // 3))
        y, x = int(rng.integers(0, h - bh)), int(rng.integers(0, w - bw))
        if occ[max(0, y - 3):y + bh + 3, max(0, x - 3):x + bw + 3].any():
            continue
        prob[y:y + bh, x:x + bw] = np.clip((0.9 if placed % 7 else 0.45) + noise * rng.standard_normal((bh, bw)), 0.0, 1.0)
        occ[y:y + bh, x:x + bw] = True
        placed += 1
    return prob


def t(name, prob):
    pp = DBPostProcessor(thresh=0.3, box_thresh=0.5, unclip_ratio=1.6)
    for _ in range(2): pp.run_device(prob, want_labels=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): buf = pp.run_device(prob, want_labels=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    recs, nc = pp.fetch(buf)
    b, h, w = prob.shape
    print(f"{name:22s} {ms:7.3f} ms  {b*h*w*8/ms/1e6:8.1f} GB/s  comps/img {nc.mean():8.1f} boxes/img {np.mean([len(r) for r in recs]):6.1f}", flush=True)

if __name__ == "__main__":
    b = 16
    rend = np.stack([render_prob_map(1024, 1024, seed=i, n_boxes=150) for i in range(4)])
    rend = torch.from_numpy(np.concatenate([rend] * 4)).cuda()
    t("rendered text", rend)
    t("all foreground", torch.full((b, 1024, 1024), 0.9, device="cuda"))
    t("all background", torch.zeros(b, 1024, 1024, device="cuda"))
    t("uniform noise", torch.rand(b, 1024, 1024, device="cuda"))
    t("noise thresh~0.7 fg", torch.rand(b, 1024, 1024, device="cuda") * 0.43)
