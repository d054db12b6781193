"""Generate the golden vectors under tests/golden/ from the REFERENCE itself (run in the authoring container).

    python tests/golden/make_golden.py

Imports /root/reference/model/det unchanged, loads the seeded state_dict of oracle/weights.py strictly (which
also proves key compatibility), runs the reference forward on CPU fp32 with forward hooks, and stores input,
outputs and per-stage taps.  /root/reference does not exist on the GPU box, so the tests only read the .npz.

Also stores (a) torchvision.ops.deform_conv2d input/output vectors for the DCN kernel (edge cases included)
and (b) the result of the cv2 half of the reference post-processing (src/det/test.py:55-106 without the
pyclipper step, which cannot run here) on rendered probability maps.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from oracle.weights import make_state_dict  # noqa: E402
from oracle import postproc_oracle as ppo  # noqa: E402


def golden_model():
    from model.det.dbnet import DBNetPP  # the reference, unchanged
    torch.manual_seed(0)
    sd = make_state_dict(seed=1234)
    model = DBNetPP(pretrained=False).eval()
    model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(99)
    x = torch.randn(1, 3, 64, 96, generator=g)
    taps = {}

    def hook(name):
        def f(_m, _i, o):
            taps[name] = o.detach().numpy().copy()
        return f

    bb = model.backbone
    for n, layer in (("c2", bb.layer1), ("c3", bb.layer2), ("c4", bb.layer3), ("c5", bb.layer4)):
        layer.register_forward_hook(hook(n))
    for li in (2, 3, 4):
        for bi, blk in enumerate(getattr(bb.model, f"layer{li}")):
            blk.conv2.register_forward_hook(hook(f"backbone.model.layer{li}.{bi}.conv2"))
    for i in range(4):
        model.neck.fpn_convs[i].register_forward_hook(hook(f"p{i + 2}"))
    model.neck.asf.register_forward_hook(hook("asf"))
    model.head.bin_conv[3].register_forward_hook(hook("head.bin_conv.3"))
    model.head.thresh_conv[3].register_forward_hook(hook("head.thresh_conv.3"))
    with torch.no_grad():
        out = model(x)
    arrays = {"x": x.numpy(), "seed": np.array(1234)}
    for k, v in out.items():
        arrays["out." + k] = v.numpy()
    for k, v in taps.items():
        arrays["tap." + k] = v
    np.savez_compressed(HERE / "dbnet_small.npz", **arrays)
    print("dbnet_small.npz:", {k: v.shape for k, v in arrays.items() if k.startswith("out.")}, len(taps), "taps")


def golden_dcn():
    import torchvision.ops as tvops
    g = torch.Generator().manual_seed(5)
    arrays = {}
    for idx, (n, c, h, w, o, s) in enumerate([(1, 8, 9, 11, 6, 1), (2, 16, 12, 10, 8, 2), (1, 4, 6, 6, 4, 1)]):
        x = torch.randn(n, c, h, w, generator=g)
        ho, wo = (h + 2 - 3) // s + 1, (w + 2 - 3) // s + 1
        off = torch.randn(n, 18, ho, wo, generator=g) * 2.5
        off[:, :, 0, 0] = -1.0   # lands exactly on h = -1 / integer coordinates
        off[:, :, -1, -1] = 1.0  # lands exactly on h = H for the last tap row
        off[:, :, ho // 2, wo // 2] = 0.0
        mask = torch.sigmoid(torch.randn(n, 9, ho, wo, generator=g))
        wt = torch.randn(o, c, 3, 3, generator=g)
        y = tvops.deform_conv2d(x, off, wt, None, stride=s, padding=1, mask=mask)
        for k, v in (("x", x), ("offset", off), ("mask", mask), ("weight", wt), ("y", y)):
            arrays[f"{idx}.{k}"] = v.numpy()
        arrays[f"{idx}.stride"] = np.array(s)
    np.savez_compressed(HERE / "dcn_torchvision.npz", **arrays)
    print("dcn_torchvision.npz written")


def golden_postproc():
    """cv2 half of DBPostProcessor on rendered maps (the part of the reference that can run here)."""
    arrays = {}
    for idx, seed in enumerate((0, 1, 2)):
        prob = ppo.render_prob_map(256, 320, seed=seed, n_boxes=24)
        rects, scores = ppo.cv2_reference_path(prob, thresh=0.3, box_thresh=0.5, min_area=10, unclip_ratio=1.6)
        arrays[f"{idx}.prob"] = prob
        arrays[f"{idx}.cv2_rects"] = np.asarray(rects, dtype=np.int64).reshape(-1, 4)
        arrays[f"{idx}.cv2_scores"] = np.asarray(scores, dtype=np.float64)
    np.savez_compressed(HERE / "postproc_cv2.npz", **arrays)
    print("postproc_cv2.npz written")


if __name__ == "__main__":
    golden_model()
    golden_dcn()
    golden_postproc()
