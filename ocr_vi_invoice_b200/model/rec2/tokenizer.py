"""Drop-in Tokenizer (reference: model/rec2/tokenizer.py:3-79): id 0 = CTC blank, id 1 = padding, ids 2.. = the sorted
unique characters.  Host string logic only (strings do not live on the GPU)."""
from __future__ import annotations

import torch


class Tokenizer:
    def __init__(self, charset):
        self.blank, self.pad = "[BLANK]", "[PAD]"
        self.specials = [self.blank, self.pad]
        self.blank_id, self.pad_id = 0, 1
        self.charset = sorted(set(charset))
        self.token_to_id = {self.blank: self.blank_id, self.pad: self.pad_id}
        for i, ch in enumerate(self.charset):
            self.token_to_id[ch] = i + len(self.specials)
        self.id_to_token = {i: t for t, i in self.token_to_id.items()}
        self.num_classes = len(self.token_to_id)

    def encode(self, texts, device="cpu"):
        """Strings -> (B, L) int64 ids, unknown characters dropped, right-padded with pad_id (L >= 1 for no texts)."""
        rows = [[self.token_to_id[c] for c in text if c in self.token_to_id] for text in texts]
        width = max((len(r) for r in rows), default=1)
        return torch.tensor([r + [self.pad_id] * (width - len(r)) for r in rows], dtype=torch.long, device=device)

    def decode(self, token_ids):
        """Id sequences -> strings; blank, padding and ids outside the table are skipped."""
        if isinstance(token_ids, torch.Tensor):
            token_ids = token_ids.tolist()
        skip = (self.blank_id, self.pad_id)
        return ["".join(self.id_to_token[i] for i in ids if i not in skip and i in self.id_to_token) for ids in token_ids]
