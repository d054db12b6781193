"""Thin host wrappers over the C-ABI operators plus the weight re-layout they expect.

Every function here either re-arranges parameters (pure torch tensor ops, run once at plan time) or
launches a kernel of libdbnet_b200.so on the current CUDA stream.  Activations are NHWC bf16.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import (EPI_CHAIN, EPI_DCN_COORDS, EPI_GELU, EPI_HEAD_DOT, EPI_OUT_F32, EPI_RELU, EPI_RES, EPI_RES_UP2, EPI_SHUFFLE2, IGEMM_CONV,
                   IGEMM_DCN, IGEMM_DCN_TMEM, IGEMM_STEM)

__all__ = [
    "Igemm", "pack_conv_weight", "pack_dcn_weight", "dcn_tmem_supported", "IGEMM_DCN_TMEM", "pack_stem_weight", "pack_convT_weight", "fold_bn", "pad_vec",
    "stem_pack", "stem_pack_u8", "resize_u8_linear", "maxpool3x3s2", "asf_fwd", "asf_combine", "AsfPlan", "db_head_fwd", "postproc", "rec_prep_boxes", "rec_prep_rects", "ctc_greedy", "svtr_stem1", "layernorm", "gelu", "mhsa", "frm_select", "decimate_rows", "log_softmax_tbc", "ref_conv2d", "ref_dcn2d",
    "EPI_RELU", "EPI_RES", "EPI_RES_UP2", "EPI_OUT_F32", "EPI_SHUFFLE2", "EPI_DCN_COORDS", "EPI_HEAD_DOT", "EPI_GELU", "EPI_CHAIN", "db_maps", "dcn_coords", "IGEMM_CONV", "IGEMM_STEM", "IGEMM_DCN",
]


# ----------------------------------------------------------------------------- parameter re-layout
def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def choose_n_pad(out_c: int) -> int:
    """GEMM N padded to a block_n the kernel supports (32/64/128 or a multiple of 256)."""
    for b in (32, 64, 128, 256):
        if out_c <= b:
            return b
    return _round_up(out_c, 256)


def pack_conv_weight(w: torch.Tensor, n_pad: Optional[int] = None) -> torch.Tensor:
    """(O, I, kh, kw) fp32 -> [n_pad][kh*kw*I] bf16 with K ordered (64-channel chunk, tap, channel in chunk).

    torchvision's deform_conv2d and nn.Conv2d flatten K as (cin, tap) (`weight.flatten(1)`).  The NHWC kernels walk
    all taps of one 64-channel chunk back to back: a (chunk, tap) K-block is one contiguous 128-byte row per pixel, the
    three taps of a kernel row share one halo row in shared memory, and the deformable gathers of a chunk re-hit the
    same neighbourhood in L1.  Input channel counts below 64 (none in DBNet++) are not supported.
    """
    o, i, kh, kw = w.shape
    assert i % 64 == 0, "input channels must be a multiple of 64"
    n_pad = n_pad or choose_n_pad(o)
    out = torch.zeros(n_pad, kh * kw * i, dtype=torch.bfloat16, device=w.device)
    t = w.reshape(o, i // 64, 64, kh * kw).permute(0, 1, 3, 2)  # (o, chunk, tap, 64)
    out[:o] = t.reshape(o, kh * kw * i).to(torch.bfloat16)
    return out.contiguous()


def dcn_k_permutation() -> torch.Tensor:
    """perm[k'] = channel (inside a 64-channel chunk) that K position k' of the tensor-memory DCN kernel holds.

    The producers of dcn2.cu hand a blended row to tcgen05.st.16x256b as four lanes x 32 bytes: lane m of the four holds
    the 16-byte pieces m and m + 4 of the row (channels 8m..8m+7 and 32+8m..32+8m+7), and the store puts its registers
    (2j, 2j+1) in TMEM columns 8j + 2m, 8j + 2m + 1, i.e. K positions 16j + 4m + (0..3).  So K position 16j + 4m + r
    carries channel 32 (j >> 1) + 8m + 4 (j & 1) + r."""
    k = torch.arange(64)
    j, m, r = k >> 4, (k >> 2) & 3, k & 3
    return 32 * (j >> 1) + 8 * m + 4 * (j & 1) + r


def pack_dcn_weight(w: torch.Tensor, n_pad: Optional[int] = None) -> torch.Tensor:
    """pack_conv_weight + the K permutation of the tensor-memory DCN kernel (IGEMM_DCN_TMEM) inside each 64-channel chunk."""
    o, i, kh, kw = w.shape
    base = pack_conv_weight(w, n_pad)  # [n_pad][chunk][tap][64]
    perm = dcn_k_permutation().to(w.device)
    return base.reshape(base.shape[0], i // 64, kh * kw, 64)[..., perm].reshape(base.shape[0], -1).contiguous()


def dcn_tmem_supported(out_c: int) -> bool:
    """Layers the tensor-memory DCN kernel takes: the whole output width is one UMMA N (ResNet-50 c3 / c4)."""
    import os
    return out_c in (128, 256) and not os.environ.get("B200_DCN_V1")


def pack_stem_weight(w: torch.Tensor, n_pad: int = 64) -> torch.Tensor:
    """(O, 3, 7, 7) -> [n_pad][4 block rows][4 blocks][16] bf16 for the space-to-depth stem.

    Window position (wy, wx) of the 8x8 input window (block row ty = wy//2, dy = wy%2, ...) carries kernel
    tap (ky, kx) = (wy-1, wx-1); the first window row/column has no tap and stays zero.
    """
    o, i, kh, kw = w.shape
    assert (i, kh, kw) == (3, 7, 7), "stem is 7x7 on 3 channels"
    full = torch.zeros(o, 3, 8, 8, dtype=torch.float32, device=w.device)
    full[:, :, 1:, 1:] = w.float()
    # (o, c, ty, dy, j, dx) -> (o, ty, j, dy, dx, c)
    t = full.reshape(o, 3, 4, 2, 4, 2).permute(0, 2, 4, 3, 5, 1).reshape(o, 4, 4, 12)
    out = torch.zeros(n_pad, 4, 4, 16, dtype=torch.bfloat16, device=w.device)
    out[:o, :, :, :12] = t.to(torch.bfloat16)
    return out.reshape(n_pad, 256).contiguous()


def pack_convT_weight(w: torch.Tensor, n_pad: Optional[int] = None) -> torch.Tensor:
    """ConvTranspose2d(k=2, s=2) weight (I, O, 2, 2) -> [4*O][I] bf16, row (dy*2+dx)*O + co."""
    i, o, kh, kw = w.shape
    assert (kh, kw) == (2, 2)
    n = 4 * o
    n_pad = n_pad or choose_n_pad(n)
    out = torch.zeros(n_pad, i, dtype=torch.bfloat16, device=w.device)
    out[:n] = w.permute(2, 3, 1, 0).reshape(n, i).to(torch.bfloat16)
    return out.contiguous()


def fold_bn(bn_weight, bn_bias, mean, var, eps: float = 1e-5, conv_bias=None):
    """Eval-mode BatchNorm2d (+ optional preceding conv bias) -> per-channel (scale, shift) fp32."""
    scale = bn_weight.float() / torch.sqrt(var.float() + eps)
    shift = bn_bias.float() - mean.float() * scale
    if conv_bias is not None:
        shift = shift + conv_bias.float() * scale
    return scale, shift


def pad_vec(v: torch.Tensor, n_pad: int, fill: float = 0.0) -> torch.Tensor:
    out = torch.full((n_pad,), fill, dtype=torch.float32, device=v.device)
    out[: v.numel()] = v.float()
    return out


# ----------------------------------------------------------------------------- operators
class Igemm:
    """One planned implicit-GEMM launch (TMA descriptors baked at creation).  Keeps its tensors alive."""

    def __init__(self, mode: int, x: torch.Tensor, w: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor,
                 y: torch.Tensor, *, batch: int, in_h: int, in_w: int, in_c: int, out_c: int, ksize: int,
                 stride: int, flags: int = 0, residual: Optional[torch.Tensor] = None,
                 offsets: Optional[torch.Tensor] = None, out_ld: Optional[int] = None, res_ld: int = 0,
                 shuf_c: int = 0, block_n: int = 0, x_ld: int = 0, stride_w: int = 0, name: str = "",
                 chain: Optional[tuple] = None, x2: Optional[torch.Tensor] = None, in_c2: int = 0, stride2: int = 1):
        """`chain` = (w2, scale2, shift2, y2, relu2): EPI_CHAIN, the next block's 1x1 conv folded into this launch.
        `x2` / `in_c2` / `stride2`: second K segment of a 1x1 conv (w = [n_pad][in_c + in_c2]), see dbnet_b200.h."""
        lib = _lib.load()
        self.name = name
        self._keep = (x, w, scale, shift, y, residual, offsets, x2)
        self._chain = chain
        n_pad = w.shape[0]
        assert scale.numel() == n_pad and shift.numel() == n_pad, "scale/shift must be padded to n_pad"
        assert w.dtype == torch.bfloat16 and w.is_contiguous()
        if out_ld is None:
            out_ld = shuf_c if (flags & EPI_SHUFFLE2) else out_c
        d = _lib.IgemmDesc(
            mode=mode, batch=batch, in_h=in_h, in_w=in_w, in_c=in_c, out_c=out_c, n_pad=n_pad, ksize=ksize,
            stride=stride, flags=flags, out_ld=out_ld, res_ld=res_ld, shuf_c=shuf_c, block_n=block_n, x_ld=x_ld,
            stride_w=stride_w,
            x=_lib.ptr(x), w=_lib.ptr(w), scale=_lib.ptr(scale), shift=_lib.ptr(shift),
            residual=_lib.ptr(residual), y=_lib.ptr(y), offsets=_lib.ptr(offsets))
        if x2 is not None:
            d.x2, d.in_c2, d.x2_ld, d.stride2 = _lib.ptr(x2), in_c2, x2.shape[-1], stride2
        if chain is not None:
            w2, s2, t2, y2, relu2 = chain
            assert w2.dtype == torch.bfloat16 and w2.is_contiguous() and w2.shape[1] == out_c
            d.flags |= EPI_CHAIN
            d.chain_w, d.chain_scale, d.chain_shift, d.chain_y = _lib.ptr(w2), _lib.ptr(s2), _lib.ptr(t2), _lib.ptr(y2)
            d.chain_n, d.chain_ld, d.chain_relu = w2.shape[0], y2.shape[-1], int(bool(relu2))
        h = C.c_void_p()
        _lib.check(lib.b200_igemm_create(C.byref(d), C.byref(h)), f"igemm_create({name})")
        self._h = h
        self._lib = lib
        self.flops = 2.0 * batch * (in_h // stride) * (in_w // (stride_w or stride)) * out_c * w.shape[1]
        if chain is not None:
            self.flops += 2.0 * batch * (in_h // stride) * (in_w // (stride_w or stride)) * chain[0].shape[0] * chain[0].shape[1]

    def info(self) -> dict:
        vals = [C.c_int32() for _ in range(7)]
        _lib.check(self._lib.b200_igemm_info(self._h, *[C.byref(v) for v in vals]))
        keys = ("out_h", "out_w", "tile_w", "tile_h", "block_n", "num_tiles", "smem_bytes")
        return {k: v.value for k, v in zip(keys, vals)}

    def __call__(self, stream: Optional[int] = None) -> None:
        st = _lib.current_stream() if stream is None else stream
        _lib.check(self._lib.b200_igemm_launch(self._h, st), f"igemm_launch({self.name})")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.b200_igemm_destroy(h)
            self._h = None


def stem_pack(img: torch.Tensor, packed: torch.Tensor) -> None:
    b, c, h, w = img.shape
    assert c == 3 and img.dtype == torch.float32 and img.is_contiguous()
    _lib.check(_lib.load().b200_stem_pack(_lib.ptr(img), _lib.ptr(packed), b, h, w, _lib.current_stream()), "stem_pack")


IMAGENET_MEAN = (0.485, 0.456, 0.406)  # src/pipeline/pipeline2.py:249-250
IMAGENET_STD = (0.229, 0.224, 0.225)


def stem_pack_u8(img_hwc: torch.Tensor, packed: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> None:
    b, h, w, c = img_hwc.shape
    assert c == 3 and img_hwc.dtype == torch.uint8 and img_hwc.is_contiguous()
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    _lib.check(_lib.load().b200_stem_pack_u8(_lib.ptr(img_hwc), _lib.ptr(packed), b, h, w, m, s, _lib.current_stream()),
               "stem_pack_u8")


def resize_u8_linear(src: torch.Tensor, dst: torch.Tensor) -> None:
    """cv2.resize(src, (W2, H2)) (INTER_LINEAR, uint8 HWC, 3 channels) on the GPU, bit-exact; dst may be a view of a batch."""
    assert src.dtype == torch.uint8 and dst.dtype == torch.uint8 and src.dim() == 3 and dst.dim() == 3
    assert src.shape[2] == 3 and dst.shape[2] == 3 and src.stride(2) == 1 and dst.stride(2) == 1
    assert src.stride(1) == 3 and dst.stride(1) == 3, "pixels must be packed RGB"
    _lib.check(_lib.load().b200_resize_u8_linear(_lib.ptr(src), src.shape[0], src.shape[1], src.stride(0), _lib.ptr(dst),
                                                 dst.shape[0], dst.shape[1], dst.stride(0), _lib.current_stream()),
               "resize_u8_linear")


def maxpool3x3s2(x: torch.Tensor, y: torch.Tensor, batch: int, h: int, w: int, c: int) -> None:
    _lib.check(_lib.load().b200_maxpool3x3s2(_lib.ptr(x), _lib.ptr(y), batch, h, w, c, _lib.current_stream()), "maxpool")


def asf_fwd(p2, p3, p4, p5, w_att, b_att, out, batch, sizes) -> None:
    (h2, w2), (h3, w3), (h4, w4), (h5, w5) = sizes
    _lib.check(_lib.load().b200_asf_fwd(_lib.ptr(p2), _lib.ptr(p3), _lib.ptr(p4), _lib.ptr(p5), _lib.ptr(w_att),
                                        _lib.ptr(b_att), _lib.ptr(out), batch, h2, w2, h3, w3, h4, w4, h5, w5,
                                        _lib.current_stream()), "asf_fwd")


def asf_combine(ps, zs, zld: int, b_att, out, batch: int, sizes) -> None:
    (h2, w2), (h3, w3), (h4, w4), (h5, w5) = sizes
    _lib.check(_lib.load().b200_asf_combine(*[_lib.ptr(t) for t in ps], *[_lib.ptr(t) for t in zs], zld, _lib.ptr(b_att),
                                            _lib.ptr(out), batch, h2, w2, h3, w3, h4, w4, h5, w5, _lib.current_stream()),
               "asf_combine")


class AsfPlan:
    """ScaleFeatureSelection blend on the tensor cores (b200_asf_plan_*): TMA descriptors baked for fixed tensors."""

    def __init__(self, ps, zs, zld: int, b_att, out, batch: int, sizes):
        lib = _lib.load()
        (h2, w2), (h3, w3), (h4, w4), (h5, w5) = sizes
        self._keep = (tuple(ps), tuple(zs), b_att, out)
        h = C.c_void_p()
        _lib.check(lib.b200_asf_plan_create(*[_lib.ptr(t) for t in ps], *[_lib.ptr(t) for t in zs], zld, _lib.ptr(b_att),
                                            _lib.ptr(out), batch, h2, w2, h3, w3, h4, w4, h5, w5, C.byref(h)), "asf_plan_create")
        self._h, self._lib = h, lib

    @staticmethod
    def supported(sizes) -> bool:
        import os
        (h2, w2), (h3, w3), (h4, w4), (h5, w5) = sizes
        return bool(_lib.load().b200_asf_plan_supported(h2, w2, h3, w3, h4, w4, h5, w5)) and not os.environ.get("B200_ASF_SIMT")

    def __call__(self) -> None:
        _lib.check(self._lib.b200_asf_plan_launch(self._h, _lib.current_stream()), "asf_plan_launch")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.b200_asf_plan_destroy(h)
            self._h = None


def dcn_coords(raw: torch.Tensor, coords: torch.Tensor, batch: int, in_h: int, in_w: int, x_ld: int, stride: int) -> None:
    """raw [b, ho, wo, 32] fp32 offset_mask_conv output -> b200_dcn_coord records [b, ho, wo, 9, 4] (4-byte words)."""
    _lib.check(_lib.load().b200_dcn_coords(_lib.ptr(raw), _lib.ptr(coords), batch, in_h, in_w, x_ld, stride,
                                           _lib.current_stream()), "dcn_coords")


def db_head_fwd(t_bin, t_thr, w2_bin, b2_bin: float, w2_thr, b2_thr: float, k: float, maps: dict, batch: int,
                h: int, w: int, ld: int = 64) -> None:
    g = lambda n: _lib.ptr(maps.get(n))
    _lib.check(_lib.load().b200_db_head_fwd(_lib.ptr(t_bin), _lib.ptr(t_thr), _lib.ptr(w2_bin), b2_bin,
                                            _lib.ptr(w2_thr), b2_thr, k, g("binary"), g("thresh"), g("thresh_binary"),
                                            g("bin_logits"), g("thresh_logits"), batch, h, w, ld, _lib.current_stream()),
               "db_head_fwd")


def db_maps(bin_logits: torch.Tensor, thr_logits: torch.Tensor, k: float, maps: dict) -> None:
    """sigmoid + DB step from the two fp32 logit maps into the requested maps (dict name -> tensor, missing = skipped)."""
    g = lambda n: _lib.ptr(maps.get(n))
    _lib.check(_lib.load().b200_db_maps(_lib.ptr(bin_logits), _lib.ptr(thr_logits), float(k), g("binary"), g("thresh"),
                                        g("thresh_binary"), g("bin_logits"), g("thresh_logits"), bin_logits.numel(),
                                        _lib.current_stream()), "db_maps")


def postproc_workspace_bytes(batch: int, h: int, w: int, max_candidates: int) -> int:
    n = _lib.load().b200_postproc_workspace_bytes(batch, h, w, max_candidates)
    if n < 0:
        raise _lib.B200Error("postproc_workspace_bytes: bad shape")
    return n


def _min_area_threshold(min_area) -> int:
    """The reference drops a box when `area < min_area` with a FLOAT min_area (pipeline2.py --det_min_area type=float,
    test.py:81); box areas here are integers, so `area >= min_area` is `area >= ceil(min_area)` (10.5 drops area 10)."""
    import math
    return max(0, int(math.ceil(float(min_area))))


def postproc(prob, labels, boxes, counts, ncomp, workspace, batch, h, w, thresh, box_thresh, unclip_ratio, min_area,
             max_candidates) -> None:
    _lib.check(_lib.load().b200_postproc(_lib.ptr(prob), _lib.ptr(labels), _lib.ptr(boxes), _lib.ptr(counts),
                                         _lib.ptr(ncomp), _lib.ptr(workspace), workspace.numel() * workspace.element_size(),
                                         batch, h, w, float(thresh), float(box_thresh), float(unclip_ratio),
                                         _min_area_threshold(min_area), int(max_candidates), _lib.current_stream()), "postproc")


def _rec_images(pages, scales):
    """ctypes array of b200_image for uint8 HWC CUDA pages (row pitch free, pixels packed RGB)."""
    if not 1 <= len(pages) <= _lib.REC_MAX_IMAGES:
        raise ValueError(f"1..{_lib.REC_MAX_IMAGES} pages per call")
    arr = (_lib.Image * len(pages))()
    for i, (pg, (sh, sw)) in enumerate(zip(pages, scales)):
        if not (pg.is_cuda and pg.dtype == torch.uint8 and pg.dim() == 3 and pg.shape[2] == 3 and pg.stride(2) == 1
                and pg.stride(1) == 3):
            raise RuntimeError("recognition crops: pages must be uint8 HWC CUDA tensors with packed RGB pixels (no CPU fallback)")
        arr[i].pixels, arr[i].h, arr[i].w, arr[i].pitch = pg.data_ptr(), pg.shape[0], pg.shape[1], pg.stride(0)
        arr[i].scale_h, arr[i].scale_w = float(sh), float(sw)
    return arr


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def rec_prep_boxes(pages, scales, boxes, counts, max_candidates: int, out, rects, total, mean=IMAGENET_MEAN,
                   std=IMAGENET_STD) -> None:
    """Detector box records (device, as b200_postproc wrote them) -> (capacity, 3, th, tw) fp32 recognition inputs.
    Replaces pipeline2.py:323-328 + crop_image (test.py:123-130) + preprocess_for_recognition (pipeline2.py:92-128)."""
    capacity, _, th, tw = out.shape
    assert out.dtype == torch.float32 and out.is_contiguous() and rects.numel() * rects.element_size() >= capacity * 24
    _lib.check(_lib.load().b200_rec_prep_boxes(_rec_images(pages, scales), len(pages), _lib.ptr(boxes), _lib.ptr(counts),
                                               int(max_candidates), _lib.ptr(out), _lib.ptr(rects), _lib.ptr(total),
                                               capacity, th, tw, _f3(mean), _f3(std), _lib.current_stream()),
               "rec_prep_boxes")


def rec_prep_rects(pages, rects_in, n: int, out, rects, total, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> None:
    """The same from explicit (image, x, y, w, h, _) int32 rectangles on the device."""
    capacity, _, th, tw = out.shape
    assert out.dtype == torch.float32 and out.is_contiguous() and rects.numel() * rects.element_size() >= capacity * 24
    _lib.check(_lib.load().b200_rec_prep_rects(_rec_images(pages, [(1.0, 1.0)] * len(pages)), len(pages), _lib.ptr(rects_in),
                                               int(n), _lib.ptr(out), _lib.ptr(rects), _lib.ptr(total), capacity, th, tw,
                                               _f3(mean), _f3(std), _lib.current_stream()),
               "rec_prep_rects")


def ctc_greedy(log_probs: torch.Tensor, blank_id: int, tokens: torch.Tensor, lengths: torch.Tensor) -> None:
    """(T, B, C) fp32 CUDA log-probabilities (class axis contiguous) -> int32 tokens [B][T] (-1 padded), lengths [B].
    Replaces the argmax + collapse loop of SVTRv2.decode_probs (model/rec2/svtrv2.py:561-571)."""
    assert log_probs.dtype == torch.float32 and log_probs.dim() == 3 and log_probs.stride(2) == 1
    T, B, C = log_probs.shape
    assert tokens.dtype == torch.int32 and tokens.is_contiguous() and tokens.numel() >= B * T and lengths.numel() >= B
    _lib.check(_lib.load().b200_ctc_greedy(_lib.ptr(log_probs), T, B, C, log_probs.stride(0), log_probs.stride(1),
                                           int(blank_id), _lib.ptr(tokens), _lib.ptr(lengths), _lib.current_stream()),
               "ctc_greedy")


# ----------------------------------------------------------------------------- SVTRv2 (model/rec2/svtrv2.py)
def svtr_stem1(x, w, scale, shift, y) -> None:
    b, c, h, wd = x.shape
    assert c == 3 and x.dtype == torch.float32 and x.is_contiguous() and y.dtype == torch.bfloat16
    _lib.check(_lib.load().b200_svtr_stem1(_lib.ptr(x), _lib.ptr(w), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(y), b, h, wd,
                                           w.shape[0], _lib.current_stream()), "svtr_stem1")


def layernorm(x, y, gamma, beta, rows: int, d: int, eps: float = 1e-5, x_ld: int = 0, y_ld: int = 0) -> None:
    _lib.check(_lib.load().b200_layernorm(_lib.ptr(x), _lib.ptr(y), _lib.ptr(gamma), _lib.ptr(beta), rows, d, x_ld or d,
                                          y_ld or d, float(eps), _lib.current_stream()), "layernorm")


def gelu(x, y=None, residual=None) -> None:
    y = x if y is None else y
    _lib.check(_lib.load().b200_gelu(_lib.ptr(x), _lib.ptr(residual), _lib.ptr(y), x.numel(), _lib.current_stream()), "gelu")


def mhsa(qkv, out, n_seq: int, n: int, d: int) -> None:
    _lib.check(_lib.load().b200_mhsa(_lib.ptr(qkv), _lib.ptr(out), n_seq, n, d, 32.0 ** -0.5, _lib.current_stream()), "mhsa")


def frm_select(q, kv, out, batch: int, h: int, w: int, d: int) -> None:
    _lib.check(_lib.load().b200_frm_select(_lib.ptr(q), _lib.ptr(kv), _lib.ptr(out), batch, h, w, d, 32.0 ** -0.5,
                                           _lib.current_stream()), "frm_select")


def decimate_rows(x, y, batch: int, in_h: int, w: int, c: int) -> None:
    _lib.check(_lib.load().b200_decimate_rows(_lib.ptr(x), _lib.ptr(y), batch, in_h, w, c, _lib.current_stream()), "decimate_rows")


def log_softmax_tbc(logits, out, batch: int, t_len: int, c: int, ld: int) -> None:
    _lib.check(_lib.load().b200_log_softmax_tbc(_lib.ptr(logits), _lib.ptr(out), batch, t_len, c, ld, _lib.current_stream()),
               "log_softmax_tbc")


def ref_conv2d(x, w, scale, shift, residual, y, batch, in_h, in_w, in_c, out_c, ksize, stride, flags, out_ld,
               res_ld=0, shuf_c=0) -> None:
    _lib.check(_lib.load().b200_ref_conv2d(_lib.ptr(x), _lib.ptr(w), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(residual),
                                           _lib.ptr(y), batch, in_h, in_w, in_c, out_c, w.shape[1], ksize, stride, flags,
                                           out_ld, res_ld, shuf_c, _lib.current_stream()), "ref_conv2d")


def ref_dcn2d(x, offsets, w, scale, shift, y, batch, in_h, in_w, in_c, out_c, stride, flags, out_ld) -> None:
    _lib.check(_lib.load().b200_ref_dcn2d(_lib.ptr(x), _lib.ptr(offsets), _lib.ptr(w), _lib.ptr(scale), _lib.ptr(shift),
                                          _lib.ptr(y), batch, in_h, in_w, in_c, out_c, stride, flags, out_ld,
                                          _lib.current_stream()), "ref_dcn2d")
