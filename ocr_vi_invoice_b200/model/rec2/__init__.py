"""Recognition side of the hot path: drop-in SVTRv2 ('base'), the tokenizer / vocabulary and the
GPU greedy CTC decode."""
from .decode import GreedyCTCDecoder, decode_probs
from .svtrv2 import SVTRv2
from .tokenizer import Tokenizer
from .vocab import VOCAB

__all__ = ["GreedyCTCDecoder", "decode_probs", "SVTRv2", "Tokenizer", "VOCAB"]
