"""GPU: greedy CTC decode (b200_ctc_greedy through the drop-in decode_probs) vs the reference goldens and the oracle."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).parent / "golden"


def test_decode_probs_equals_reference_strings():
    from ocr_vi_invoice_b200.model.rec2 import VOCAB, GreedyCTCDecoder, Tokenizer, decode_probs
    g = np.load(GOLD / "ctc_reference.npz")
    tok = Tokenizer(VOCAB)
    dec = GreedyCTCDecoder(tok)
    for name in "abc":
        lp = torch.from_numpy(g[f"{name}_log_probs"]).cuda()
        assert decode_probs(lp, tok) == g[f"{name}_texts"].tolist()
        tokens, lengths = dec.run_device(lp)
        assert np.array_equal(lengths.cpu().numpy(), g[f"{name}_lengths"])
        assert np.array_equal(tokens.cpu().numpy(), g[f"{name}_ids"])  # ids then -1 padding, bit for bit


def test_random_shapes_and_strides_against_oracle():
    from oracle import ctc_oracle as co
    from ocr_vi_invoice_b200.model.rec2 import VOCAB, GreedyCTCDecoder, Tokenizer
    rng = np.random.default_rng(3)
    dec = GreedyCTCDecoder(Tokenizer(VOCAB))
    for T, B, C in ((1, 1, 1), (7, 3, 2), (64, 32, 232), (300, 5, 33), (1000, 2, 500)):
        lp = (np.round(rng.normal(0, 1, size=(T, B, C)) * 2) / 2).astype(np.float32)  # many exact ties
        lp[rng.random((T, B)) < 0.5, 0] = 9.0                                          # many blanks
        if C > 2:
            lp[rng.random((T, B)) < 0.02, 2] = np.nan
            lp[rng.random((T, B)) < 0.02, :] = -np.inf                                 # all -inf -> class 0
        want = co.greedy_ids(lp)
        assert dec.ids(torch.from_numpy(lp).cuda()) == want
        # the layout the recogniser really produces: (B, T, C) logits permuted to (T, B, C) (svtrv2.py:536-537)
        bt = torch.from_numpy(np.ascontiguousarray(lp.transpose(1, 0, 2))).cuda()
        assert dec.ids(bt.permute(1, 0, 2)) == want
