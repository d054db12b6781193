"""GPU greedy CTC decode with the reference's interface (SVTRv2.decode_probs, model/rec2/svtrv2.py:550-569).

`decode_probs(log_probs, tokenizer)` takes the (T, B, num_classes) CUDA tensor the recogniser returns and gives the list
of strings; the argmax + repeat-collapse + blank removal run in one kernel (`b200_ctc_greedy`), B * (T + 1) int32 come
back, and only the id -> character join stays on the host.  No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

from typing import List

import torch

from ... import ops
from .tokenizer import Tokenizer


class GreedyCTCDecoder:
    def __init__(self, tokenizer: Tokenizer):
        self.tokenizer = tokenizer
        self.blank_id = tokenizer.blank_id
        self._buf = None

    def run_device(self, log_probs: torch.Tensor):
        """Queues the kernel on the current stream; returns device (tokens [B][T] int32, -1 padded; lengths [B])."""
        if not (isinstance(log_probs, torch.Tensor) and log_probs.is_cuda and log_probs.dim() == 3):
            raise RuntimeError("decode_probs (B200): expects a (T, B, num_classes) CUDA tensor (no CPU fallback)")
        lp = log_probs.float()
        if lp.stride(2) != 1:
            lp = lp.contiguous()
        T, B, _ = lp.shape
        if self._buf is None or self._buf[0].shape != (B, T) or self._buf[0].device != lp.device:
            self._buf = (torch.empty(B, T, dtype=torch.int32, device=lp.device),
                         torch.empty(B, dtype=torch.int32, device=lp.device))
        with torch.cuda.device(lp.device):
            ops.ctc_greedy(lp, self.blank_id, self._buf[0], self._buf[1])
        return self._buf

    def ids(self, log_probs: torch.Tensor) -> List[List[int]]:
        tokens, lengths = self.run_device(log_probs)
        tokens, lengths = tokens.cpu().numpy(), lengths.cpu().numpy()
        return [tokens[b, :lengths[b]].tolist() for b in range(len(lengths))]

    def __call__(self, log_probs: torch.Tensor) -> List[str]:
        return self.tokenizer.decode(self.ids(log_probs))


def decode_probs(log_probs: torch.Tensor, tokenizer: Tokenizer) -> List[str]:
    return GreedyCTCDecoder(tokenizer)(log_probs)
