"""ctypes binding of the C-ABI in include/dbnet_b200.h (libdbnet_b200.so).

The reference is pure Python and reaches its kernels through torch / torchvision / cv2 calls
(model/det/*.py, src/det/test.py:55-106); this module is the one place where the drop-in reaches
the hand-written sm_100a kernels instead.  There is no fallback: if the library is missing or a
call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libdbnet_b200.so"

IGEMM_CONV, IGEMM_STEM, IGEMM_DCN, IGEMM_DCN_TMEM = 0, 1, 2, 3
EPI_RELU, EPI_RES, EPI_RES_UP2, EPI_OUT_F32, EPI_SHUFFLE2, EPI_DCN_COORDS, EPI_HEAD_DOT, EPI_GELU, EPI_CHAIN = 1, 2, 4, 8, 16, 32, 64, 128, 256


class B200Error(RuntimeError):
    pass


class IgemmDesc(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("batch", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32),
        ("in_c", C.c_int32), ("out_c", C.c_int32), ("n_pad", C.c_int32), ("ksize", C.c_int32),
        ("stride", C.c_int32), ("flags", C.c_int32), ("out_ld", C.c_int32), ("res_ld", C.c_int32),
        ("shuf_c", C.c_int32), ("block_n", C.c_int32), ("x_ld", C.c_int32), ("stride_w", C.c_int32),
        ("x", C.c_void_p), ("w", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("residual", C.c_void_p), ("y", C.c_void_p), ("offsets", C.c_void_p),
        ("chain_w", C.c_void_p), ("chain_scale", C.c_void_p), ("chain_shift", C.c_void_p), ("chain_y", C.c_void_p),
        ("chain_n", C.c_int32), ("chain_ld", C.c_int32), ("chain_relu", C.c_int32), ("chain_reserved", C.c_int32),
        ("x2", C.c_void_p), ("in_c2", C.c_int32), ("x2_ld", C.c_int32), ("stride2", C.c_int32), ("x2_reserved", C.c_int32),
    ]


class Box(C.Structure):
    _fields_ = [
        ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
        ("bx0", C.c_int32), ("by0", C.c_int32), ("bx1", C.c_int32), ("by1", C.c_int32),
        ("label", C.c_int32), ("npix", C.c_int32), ("qsum", C.c_uint64),
    ]


class Image(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("pitch", C.c_int32), ("reserved", C.c_int32),
                ("scale_h", C.c_double), ("scale_w", C.c_double)]


class RecRect(C.Structure):
    _fields_ = [("image", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32),
                ("status", C.c_int32)]


REC_OK, REC_EMPTY, REC_DEGENERATE = 0, 1, 2
REC_MAX_IMAGES = 32

# every symbol include/dbnet_b200.h declares: name -> (restype, argtypes)
_i, _p, _f, _d, _i64 = C.c_int, C.c_void_p, C.c_float, C.c_double, C.c_int64
SYMBOLS = {
    "b200_abi_version": (_i, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_device_check": (_i, [_i]),
    "b200_igemm_create": (_i, [C.POINTER(IgemmDesc), C.POINTER(_p)]),
    "b200_igemm_launch": (_i, [_p, _p]),
    "b200_igemm_destroy": (_i, [_p]),
    "b200_igemm_info": (_i, [_p] + [C.POINTER(C.c_int32)] * 7),
    "b200_stem_pack": (_i, [_p, _p, _i, _i, _i, _p]),
    "b200_stem_pack_u8": (_i, [_p, _p, _i, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float), _p]),
    "b200_db_maps": (_i, [_p, _p, _f, _p, _p, _p, _p, _p, _i64, _p]),
    "b200_resize_u8_linear": (_i, [_p, _i, _i, _i, _p, _i, _i, _i, _p]),
    "b200_maxpool3x3s2": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "b200_asf_fwd": (_i, [_p] * 7 + [_i] * 9 + [_p]),
    "b200_asf_combine": (_i, [_p] * 8 + [_i] + [_p, _p] + [_i] * 9 + [_p]),
    "b200_asf_plan_supported": (_i, [_i] * 8),
    "b200_asf_plan_create": (_i, [_p] * 8 + [_i] + [_p, _p] + [_i] * 9 + [C.POINTER(_p)]),
    "b200_asf_plan_launch": (_i, [_p, _p]),
    "b200_asf_plan_destroy": (_i, [_p]),
    "b200_db_head_fwd": (_i, [_p, _p, _p, _f, _p, _f, _f, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "b200_dcn_coords": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "b200_postproc_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "b200_postproc": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _d, _d, _d, _i, _i, _p]),
    "b200_rec_prep_boxes": (_i, [C.POINTER(Image), _i, _p, _p, _i, _p, _p, _p, _i, _i, _i, C.POINTER(C.c_float),
                                 C.POINTER(C.c_float), _p]),
    "b200_rec_prep_rects": (_i, [C.POINTER(Image), _i, _p, _i, _p, _p, _p, _i, _i, _i, C.POINTER(C.c_float),
                                 C.POINTER(C.c_float), _p]),
    "b200_ctc_greedy": (_i, [_p, _i, _i, _i, _i64, _i64, _i, _p, _p, _p]),
    "b200_svtr_stem1": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "b200_layernorm": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i, _f, _p]),
    "b200_gelu": (_i, [_p, _p, _p, _i64, _p]),
    "b200_mhsa": (_i, [_p, _p, _i, _i, _i, _f, _p]),
    "b200_frm_select": (_i, [_p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "b200_decimate_rows": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "b200_log_softmax_tbc": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "b200_ref_conv2d": (_i, [_p] * 6 + [_i] * 12 + [_p]),
    "b200_ref_dcn2d": (_i, [_p] * 6 + [_i] * 8 + [_p]),
}

_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load libdbnet_b200.so (built in-tree by ocr_vi_invoice_b200.build); fail loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise B200Error(
            f"{_LIB_PATH} is missing: build it with `python -m ocr_vi_invoice_b200.build` "
            "(there is no CPU or torch fallback for the DBNet++ path)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.b200_abi_version() != 3:
        raise B200Error("libdbnet_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().b200_last_error().decode(errors="replace")
        raise B200Error(f"{what}: {msg}" if what else msg)


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
