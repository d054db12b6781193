"""Golden vectors for the greedy CTC decode from the REFERENCE'S OWN CODE (authoring container only).

    python tests/golden/make_ctc_golden.py

Imports /root/reference/model/rec2/svtrv2.py unchanged, builds `SVTRv2(variant='tiny')` (random weights are irrelevant:
only `decode_probs`, `blank_id` and the tokenizer are used) and runs `model.decode_probs(log_probs)`
(svtrv2.py:550-569) on seeded (T, B, C) tensors built to hit: exact ties between classes (values quantised to 1/4),
runs of repeated characters with and without blanks between them, all-blank sequences, class 1 ([PAD]) winning, and NaN
entries.  Stored: the tensors, the decoded strings, the id lists (re-derived with the reference's loop) and the charset.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, "/root/reference")


def main():
    from model.rec2.svtrv2 import SVTRv2
    from model.rec2.vocab import VOCAB
    model = SVTRv2(variant="tiny").eval()
    C = model.tokenizer.num_classes
    rng = np.random.default_rng(21)
    out = {"charset": np.array(list(VOCAB)), "num_classes": np.int64(C)}
    for name, (T, B) in {"a": (64, 24), "b": (40, 7), "c": (128, 5)}.items():
        lp = np.round(rng.normal(0, 2, size=(T, B, C)) * 4) / 4          # quantised: exact ties are common
        path = rng.integers(0, C, size=(T, B))
        path[rng.random((T, B)) < 0.45] = 0                                # blanks
        rep = rng.random((T, B)) < 0.35                                    # repeats of the previous frame
        for t in range(1, T):
            path[t, rep[t]] = path[t - 1, rep[t]]
        strong = rng.random((T, B)) < 0.8                                  # most frames: a clear winner
        tt, bb = np.nonzero(strong)
        lp[tt, bb, path[tt, bb]] = 30.0
        lp[:, 0, :] = -5.0
        lp[:, 0, 0] = 0.0                                                   # sequence 0: all blank
        lp[3, 1, :] = 1.0                                                   # a frame where every class ties -> index 0
        lp[5, 1, 1] = 50.0                                                  # [PAD] wins a frame
        lp[7, 2, 40] = np.nan                                               # NaN counts as the maximum
        lp[7, 2, 90] = np.nan
        lp = lp.astype(np.float32)
        t = torch.from_numpy(lp)
        texts = model.decode_probs(t)
        # the reference's loop once more, to keep the ids as well (svtrv2.py:561-571)
        preds = t.argmax(dim=-1).permute(1, 0)
        ids = []
        for pred in preds:
            chars, prev = [], None
            for p in pred.tolist():
                if p != model.blank_id and p != prev:
                    chars.append(p)
                prev = p
            ids.append(chars)
        assert model.tokenizer.decode(ids) == texts
        out[f"{name}_log_probs"] = lp
        out[f"{name}_texts"] = np.array(texts)
        out[f"{name}_lengths"] = np.array([len(i) for i in ids], dtype=np.int32)
        padded = np.full((B, T), -1, dtype=np.int32)
        for b, i in enumerate(ids):
            padded[b, :len(i)] = i
        out[f"{name}_ids"] = padded
        print(name, lp.shape, "mean length", np.mean([len(i) for i in ids]), repr(texts[1][:30]))
    np.savez_compressed(HERE / "ctc_reference.npz", **out)
    print("wrote", HERE / "ctc_reference.npz")


if __name__ == "__main__":
    main()
