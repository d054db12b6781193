"""CPU restatement of the recogniser's greedy CTC decode -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

SURVEY.md 8(f) item 3, the tail of the recognition path.  Reference (file:line under the reference tree):
  * `SVTRv2.decode_probs`      model/rec2/svtrv2.py:550-569  (argmax over classes, collapse repeats, drop blanks)
  * `Tokenizer.__init__/decode` model/rec2/tokenizer.py:3-22, 54-79 (ids: 0 blank, 1 pad, 2.. sorted charset; decode
    skips blank / pad / unknown ids and joins)
  * `VOCAB`                    model/rec2/vocab.py:1-21 (the charset is passed in here, not restated: tests use the
    charset stored in the golden file)

Third-party arithmetic: `torch.argmax` (unpinned torch; 2.11 here) -- the FIRST index of the maximum, a NaN counts as
the maximum.  PINNED: tests/golden/ctc_reference.npz holds `decode_probs` outputs of the reference's own SVTRv2 class
(imported unchanged by tests/golden/make_ctc_golden.py) on seeded log-probabilities with exact ties, long repeats,
all-blank rows and NaNs; this restatement matches them (tests/test_oracle.py).
"""
from __future__ import annotations

import numpy as np

BLANK_ID = 0
PAD_ID = 1


def argmax_first(row: np.ndarray) -> int:
    """torch.argmax over one row: first index of the maximum; the first NaN wins if there is one."""
    nan = np.isnan(row)
    if nan.any():
        return int(np.flatnonzero(nan)[0])
    return int(np.flatnonzero(row == row.max())[0])


def greedy_ids(log_probs: np.ndarray, blank_id: int = BLANK_ID):
    """(T, B, C) float -> list of B id lists (svtrv2.py:561-571): keep p iff p != blank and p != previous frame's p."""
    T, B, _ = log_probs.shape
    out = []
    for b in range(B):
        ids, prev = [], None
        for t in range(T):
            p = argmax_first(log_probs[t, b])
            if p != blank_id and p != prev:
                ids.append(p)
            prev = p
        out.append(ids)
    return out


def make_id_to_token(charset):
    """tokenizer.py:16-20."""
    chars = list(sorted(set(charset)))
    table = {BLANK_ID: "[BLANK]", PAD_ID: "[PAD]"}
    table.update({i + 2: t for i, t in enumerate(chars)})
    return table


def decode_ids(id_lists, id_to_token):
    """tokenizer.py:64-79."""
    return ["".join(id_to_token[i] for i in ids if i not in (BLANK_ID, PAD_ID) and i in id_to_token) for ids in id_lists]


def decode_probs(log_probs: np.ndarray, charset) -> list:
    return decode_ids(greedy_ids(log_probs), make_id_to_token(charset))
