"""Times the GPU post-processing on batch-16 1024x1024 maps of different character (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ocr_vi_invoice_b200.postprocess import DBPostProcessor
from oracle import postproc_oracle as ppo

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
    rend = np.stack([ppo.render_prob_map(1024, 1024, seed=i, n_boxes=150) for i in range(4)])
    rend = torch.from_numpy(np.concatenate([rend] * 4)).cuda()
    t("rendered text", rend)
    t("all foreground", torch.full((b, 1024, 1024), 0.9, device="cuda"))
    t("all background", torch.zeros(b, 1024, 1024, device="cuda"))
    t("uniform noise", torch.rand(b, 1024, 1024, device="cuda"))
    t("noise thresh~0.7 fg", torch.rand(b, 1024, 1024, device="cuda") * 0.43)
