"""The recogniser's character set (reference: model/rec2/vocab.py:1-21), built from its description instead of a typed
string: the twelve Vietnamese vowels with each of the six tones in both cases, every ASCII letter plus d-with-stroke,
the digits, ASCII punctuation, the space and the dong sign.  tests/test_host.py checks it against the charset the
reference's tokenizer was built with (stored in tests/golden/ctc_reference.npz)."""
import string
import unicodedata

_BASE_VOWELS = "aăâeêioôơuưy"          # a ă â e ê i o ô ơ u ư y
_TONES = ("", "́", "̀", "̉", "̃", "̣")       # level, acute, grave, hook above, tilde, dot below


def _build() -> str:
    vowels = [unicodedata.normalize("NFC", v + t) for v in _BASE_VOWELS for t in _TONES]
    chars = set(vowels) | {v.upper() for v in vowels}
    chars |= set(string.ascii_letters) | {"đ", "Đ"}         # đ Đ
    chars |= set(string.digits) | set(string.punctuation) | {" ", "₫"}  # ₫
    return "".join(sorted(chars))


VOCAB = _build()
