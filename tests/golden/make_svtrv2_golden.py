"""Golden vectors for the SVTRv2 recogniser from the REFERENCE'S OWN MODULE (authoring container only).

    python tests/golden/make_svtrv2_golden.py

Imports /root/reference/model/rec2/svtrv2.py unchanged, builds `SVTRv2(variant='base').eval()`, loads the seeded
parameters of oracle.svtrv2_oracle.make_state_dict STRICTLY (so the key set and shapes are checked against the real
module), and runs its forward on three 32 x 256 inputs produced the way the pipeline produces them (uint8 crops through
the reference-pinned preprocess_for_recognition restatement: one text-like crop, one noise crop, one narrow crop with
white padding).  Stored: the inputs, the (T, B, C) log-probabilities, the decoded strings, and stage taps captured with
forward hooks on the reference's sub-modules (every 8th token, to keep the fixture small).
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from oracle import recprep_oracle as ro  # noqa: E402
from oracle import svtrv2_oracle as so  # noqa: E402

SEED = 5
TAP_STRIDE = 8


def inputs():
    rng = np.random.default_rng(17)
    crops = []
    text = np.full((40, 300, 3), 240, dtype=np.uint8)
    for x in range(8, 290, 14):
        text[8:32, x:x + int(rng.integers(4, 11))] = rng.integers(0, 90, size=3)
    crops.append(text)
    crops.append(rng.integers(0, 256, size=(25, 410, 3), dtype=np.uint8))
    narrow = np.full((48, 90, 3), 225, dtype=np.uint8)
    narrow[10:40, 10:80:9] = 30
    crops.append(narrow)
    return np.stack([ro.preprocess_for_recognition(c)[0] for c in crops])


def main():
    from model.rec2.svtrv2 import SVTRv2
    torch.manual_seed(0)
    model = SVTRv2(variant="base").eval()
    sd = so.make_state_dict("base", SEED, model.tokenizer.num_classes)
    assert list(sd.keys()) == list(model.state_dict().keys()), "key order / set differs from the reference module"
    model.load_state_dict(sd, strict=True)
    x = torch.from_numpy(inputs())
    taps = {}
    hooks = [model.stem.register_forward_hook(lambda m, i, o: taps.__setitem__("stem", o.flatten(2).transpose(1, 2))),
             model.frm.register_forward_hook(lambda m, i, o: taps.__setitem__("frm", o)),
             model.head.register_forward_hook(lambda m, i, o: taps.__setitem__("logits", o))]
    for s in range(3):
        hooks.append(model.stages[s].register_forward_hook(lambda m, i, o, s=s: taps.__setitem__(f"stage{s}", o)))
    with torch.no_grad():
        lp = model(x)
        texts = model.decode_probs(lp)
    for h in hooks:
        h.remove()
    out = {"seed": np.int64(SEED), "x": x.numpy(), "log_probs": lp.numpy(), "texts": np.array(texts),
           "tap_stride": np.int64(TAP_STRIDE)}
    for k, v in taps.items():
        out["tap." + k] = v.numpy()[:, ::TAP_STRIDE] if k.startswith("sta") or k == "stem" else v.numpy()
        print(k, tuple(v.shape), float(v.abs().mean()))
    # the other two variants (the class default is 'small'): log-probabilities on the same inputs
    for variant in ("tiny", "small"):
        mv = SVTRv2(variant=variant).eval()
        sdv = so.make_state_dict(variant, SEED, mv.tokenizer.num_classes)
        assert list(sdv.keys()) == list(mv.state_dict().keys())
        mv.load_state_dict(sdv, strict=True)
        with torch.no_grad():
            out[f"log_probs_{variant}"] = mv(x).numpy()
    print("log_probs", lp.shape, "max prob per frame (mean)", float(lp.exp().max(-1).values.mean()), texts)
    np.savez_compressed(HERE / "svtrv2_small.npz", **out)
    print("wrote", HERE / "svtrv2_small.npz")


if __name__ == "__main__":
    main()
