"""Recognition-side pieces of the hot path that exist so far: the tokenizer and the GPU greedy
CTC decode.  The SVTRv2 network itself is not built yet."""
from .decode import GreedyCTCDecoder, decode_probs
from .tokenizer import Tokenizer
from .vocab import VOCAB

__all__ = ["GreedyCTCDecoder", "decode_probs", "Tokenizer", "VOCAB"]
