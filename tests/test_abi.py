"""CPU: the C-ABI library builds, loads without a GPU, and exports exactly what include/dbnet_b200.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from ocr_vi_invoice_b200 import build, _lib
    build.build_library()
    return _lib.load()


def _declared():
    text = (ROOT / "include" / "dbnet_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from ocr_vi_invoice_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in dbnet_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"


def test_status_calls_work_without_gpu(lib):
    assert lib.b200_abi_version() == 3
    assert lib.b200_postproc_workspace_bytes(1, 64, 64, 100) > 2 * 64 * 64 * 4
    assert lib.b200_postproc_workspace_bytes(0, 64, 64, 100) == -1
    # argument validation happens before any CUDA call and reports through b200_last_error
    h = ctypes.c_void_p()
    assert lib.b200_igemm_create(None, ctypes.byref(h)) != 0
    assert b"null" in lib.b200_last_error()


def test_box_record_layout():
    from ocr_vi_invoice_b200 import _lib, postprocess
    assert ctypes.sizeof(_lib.Box) == 48 == postprocess._BOX_DTYPE.itemsize
    assert _lib.Box.qsum.offset == 40 == postprocess._BOX_DTYPE.fields["qsum"][1]
    assert ctypes.sizeof(_lib.IgemmDesc) == 16 * 4 + 12 * 8 + 8 * 4  # ABI 2: chain_* fields appended; ABI 3: the second K segment (x2)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from ocr_vi_invoice_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.B200Error, match="no CPU or torch fallback"):
        _lib.load()
