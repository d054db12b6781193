"""GPU parity of the tcgen05 implicit-GEMM family (conv / stem / convT / DCN) through the C-ABI.

Each case is checked against (a) torch fp32 ops (TF32 off) on the same bf16-rounded operands and, for the
conv/DCN cases, (b) the library's own CUDA-core cross-check kernels.  Tolerances: outputs are bf16
(rel 2^-8) or fp32 with fp32 accumulation in a different order.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def nhwc(x):  # NCHW fp32 -> NHWC bf16
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(y):  # NHWC any -> NCHW fp32
    return y.float().permute(0, 3, 1, 2).contiguous()


def close(ours, ref, rel_rms=2e-2, rtol=1.6e-2):
    rms = ref.float().pow(2).mean().sqrt().item() + 1e-12
    err = (ours.float() - ref.float()).abs()
    bound = rel_rms * rms + rtol * ref.float().abs()
    bad = (err > bound)
    if bad.numel() > (1 << 22) and bad.sum().item() <= bad.numel() // (1 << 22) and err.max().item() < 2.5 * bound.min().item():
        return  # millions of elements: the far tail of the bf16 rounding noise pokes through a fixed bound (< 1 in 4M)
    assert not bad.any(), (
        f"mismatch: {bad.sum().item()}/{bad.numel()} elements, max err {err.max().item():.4g}, rms {rms:.4g}, "
        f"first bad index {bad.nonzero()[0].tolist()}")


CONV_CASES = [
    # (batch, H, W, Cin, Cout, k, stride, flags-tag)
    (2, 32, 32, 64, 64, 1, 1, ""),
    (1, 16, 128, 64, 256, 1, 1, "relu"),
    (2, 20, 20, 256, 64, 1, 1, "relu"),
    (1, 24, 40, 128, 128, 3, 1, "relu"),
    (2, 16, 16, 256, 256, 3, 1, ""),
    (1, 8, 8, 512, 512, 1, 1, "res+relu"),
    (1, 30, 22, 64, 64, 3, 1, "res+relu"),
    (1, 32, 32, 256, 512, 1, 2, ""),
    (2, 20, 36, 128, 128, 3, 2, "relu"),
    (1, 64, 64, 128, 27, 3, 1, "f32"),
    (2, 40, 24, 256, 27, 3, 2, "f32"),
    (1, 16, 16, 2048, 256, 1, 1, "bias"),
    (1, 16, 16, 512, 256, 1, 1, "up2"),
    (1, 64, 256, 64, 64, 3, 1, "relu"),
    (3, 12, 12, 1024, 2048, 1, 1, ""),
    (2, 48, 40, 64, 256, 1, 1, "res+relu"),
    (1, 36, 28, 128, 512, 1, 1, "res+relu"),
    (1, 64, 48, 256, 256, 1, 1, "up2"),
    (2, 10, 6, 256, 1024, 1, 1, "res+relu"),
    (1, 128, 128, 64, 64, 1, 1, "relu"),
    (1, 6, 128, 64, 64, 3, 1, "relu"),          # 128x1 tiles
    (2, 5, 200, 256, 256, 3, 1, ""),            # ragged second tile, 4 channel chunks
    (1, 9, 256, 128, 27, 3, 1, "f32"),          # 128x1 tiles with the direct fp32 epilogue
    (1, 4, 130, 256, 128, 3, 1, "res+relu"),    # 128x1 tiles + residual
    (2, 40, 56, 256, 128, 3, 1, "relu"),        # share mode on a CTA pair (DB head 256 -> 128), ragged right edge
    (3, 24, 24, 128, 128, 3, 1, "res+relu"),    # the same with a residual and an odd tile count
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "b%d_%dx%d_c%d_o%d_k%d_s%d_%s" % c)
def test_conv_igemm(case):
    from ocr_vi_invoice_b200 import ops
    _setup()
    b, h, w, ci, co, k, s, tag = case
    g = torch.Generator(device="cuda").manual_seed(sum(int(v * 7) for v in case if not isinstance(v, str)) & 0xffff)
    dev = "cuda"
    x = torch.randn(b, ci, h, w, device=dev, generator=g)
    wt = torch.randn(co, ci, k, k, device=dev, generator=g) / (ci * k * k) ** 0.5
    scale = torch.rand(co, device=dev, generator=g) + 0.5
    shift = torch.randn(co, device=dev, generator=g) * 0.1
    xb, wb = x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float()
    oh, ow = h // s, w // s
    flags = 0
    res = None
    ref = F.conv2d(xb, wb, stride=s, padding=(k - 1) // 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    res_ld = 0
    if "res" in tag:
        r = torch.randn(b, co, oh, ow, device=dev, generator=g)
        res = nhwc(r)
        ref = ref + res.float().permute(0, 3, 1, 2)
        flags |= ops.EPI_RES
        res_ld = co
    if "up2" in tag:
        r = torch.randn(b, co, oh // 2, ow // 2, device=dev, generator=g)
        res = nhwc(r)
        ref = ref + F.interpolate(res.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
        flags |= ops.EPI_RES_UP2
        res_ld = co
    if "relu" in tag:
        ref = ref.relu()
        flags |= ops.EPI_RELU
    out_c = co
    n_pad = ops.choose_n_pad(co)
    if "f32" in tag:
        flags |= ops.EPI_OUT_F32
        out_c = 32
        y = torch.full((b, oh, ow, 32), float("nan"), device=dev, dtype=torch.float32)
    else:
        y = torch.full((b, oh, ow, co), float("nan"), device=dev, dtype=torch.bfloat16)
    wp = ops.pack_conv_weight(wt, n_pad)
    sc, sh = ops.pad_vec(scale, n_pad, 1.0), ops.pad_vec(shift, n_pad)
    op = ops.Igemm(ops.IGEMM_CONV, nhwc(x), wp, sc, sh, y, batch=b, in_h=h, in_w=w, in_c=ci, out_c=out_c, ksize=k,
                   stride=s, flags=flags, residual=res, res_ld=res_ld, name=str(case))
    op()
    torch.cuda.synchronize()
    got = nchw(y)[:, :co]
    if "f32" in tag:
        close(got, ref, rel_rms=1e-3, rtol=1e-3)
        assert (nchw(y)[:, co:] == 0).all()
    else:
        close(got, ref)
    # cross-check kernel, same flags
    y2 = torch.zeros_like(y)
    ops.ref_conv2d(nhwc(x), wp, sc, sh, res, y2, b, h, w, ci, out_c, k, s, flags, y.shape[-1], res_ld)
    torch.cuda.synchronize()
    close(nchw(y2)[:, :co], ref, rel_rms=(1e-3 if "f32" in tag else 2e-2))


@pytest.mark.parametrize("shape", [(1, 64, 64), (2, 96, 160), (1, 256, 224)])
def test_stem_igemm(shape):
    from ocr_vi_invoice_b200 import ops
    _setup()
    b, h, w = shape
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(7)
    img = torch.randn(b, 3, h, w, device=dev, generator=g)
    wt = torch.randn(64, 3, 7, 7, device=dev, generator=g) / 147 ** 0.5
    scale = torch.rand(64, device=dev, generator=g) + 0.5
    shift = torch.randn(64, device=dev, generator=g) * 0.1
    packed = torch.empty(b, h // 2 + 3, w // 2 + 3, 16, device=dev, dtype=torch.bfloat16)
    ops.stem_pack(img, packed)
    y = torch.full((b, h // 2, w // 2, 64), float("nan"), device=dev, dtype=torch.bfloat16)
    op = ops.Igemm(ops.IGEMM_STEM, packed, ops.pack_stem_weight(wt, 64), scale, shift, y, batch=b, in_h=h, in_w=w,
                   in_c=3, out_c=64, ksize=7, stride=2, flags=ops.EPI_RELU, name="stem")
    op()
    torch.cuda.synchronize()
    ref = F.conv2d(img.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), stride=2, padding=3)
    ref = (ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).relu()
    close(nchw(y), ref)
    # max-pool on top
    yp = torch.empty(b, h // 4, w // 4, 64, device=dev, dtype=torch.bfloat16)
    ops.maxpool3x3s2(y, yp, b, h // 2, w // 2, 64)
    torch.cuda.synchronize()
    refp = F.max_pool2d(nchw(y), 3, 2, 1)
    assert torch.equal(nchw(yp), refp)


@pytest.mark.parametrize("case", [(2, 20, 36, 8), (1, 34, 70, 32), (3, 12, 18, 128), (1, 64, 66, 64), (1, 8, 10, 24)])
def test_maxpool_channel_counts(case):
    """The strip kernel's block shapes (256 / (C / 8) output columns per block) and the flat-index fallback (C / 8 = 3)."""
    from ocr_vi_invoice_b200 import ops
    _setup()
    b, h, w, c = case
    g = torch.Generator(device="cuda").manual_seed(c)
    x = torch.randn(b, h, w, c, device="cuda", generator=g).to(torch.bfloat16)
    y = torch.full((b, h // 2, w // 2, c), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.maxpool3x3s2(x, y, b, h, w, c)
    torch.cuda.synchronize()
    assert torch.equal(nchw(y), F.max_pool2d(nchw(x), 3, 2, 1))


@pytest.mark.parametrize("shape", [(1, 16, 16), (2, 24, 40)])
def test_convT_igemm(shape):
    from ocr_vi_invoice_b200 import ops
    _setup()
    b, h, w = shape
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(11)
    x = torch.randn(b, 64, h, w, device=dev, generator=g)
    wt = torch.randn(64, 64, 2, 2, device=dev, generator=g) / 8
    bias = torch.randn(64, device=dev, generator=g) * 0.1
    scale = torch.rand(64, device=dev, generator=g) + 0.5
    shift = torch.randn(64, device=dev, generator=g) * 0.1
    ref = F.conv_transpose2d(x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), bias, stride=2)
    ref = (ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).relu()
    sc4 = scale.repeat(4)
    sh4 = (shift + bias * scale).repeat(4)
    y = torch.full((b, 2 * h, 2 * w, 64), float("nan"), device=dev, dtype=torch.bfloat16)
    op = ops.Igemm(ops.IGEMM_CONV, nhwc(x), ops.pack_convT_weight(wt, 256), sc4, sh4, y, batch=b, in_h=h, in_w=w,
                   in_c=64, out_c=256, ksize=1, stride=1, flags=ops.EPI_RELU | ops.EPI_SHUFFLE2, shuf_c=64,
                   name="convT")
    op()
    torch.cuda.synchronize()
    close(nchw(y), ref)


@pytest.mark.parametrize("n2", [64, 128])
@pytest.mark.parametrize("shape", [(1, 16, 24), (2, 40, 56), (16, 128, 128)], ids=lambda c: "b%d_%dx%d" % c)
def test_chained_second_gemm(shape, n2):
    """B200_EPI_CHAIN: conv3 (64 -> 256) + residual + ReLU of one Bottleneck with the next block's conv1 (256 -> 64) + BN + ReLU
    chained onto its staged output chunks -- y as usual, y2 == the 1x1 conv applied to the bf16 y the launch stored."""
    from ocr_vi_invoice_b200 import ops
    _setup()
    b, h, w = shape
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(b * 1000 + h)
    x = torch.randn(b, 64, h, w, device=dev, generator=g)
    r = torch.randn(b, 256, h, w, device=dev, generator=g)
    wt = torch.randn(256, 64, 1, 1, device=dev, generator=g) / 8
    w2 = torch.randn(n2, 256, 1, 1, device=dev, generator=g) / 16
    sc, sh = torch.rand(256, device=dev, generator=g) + 0.5, torch.randn(256, device=dev, generator=g) * 0.1
    s2, t2 = torch.rand(n2, device=dev, generator=g) + 0.5, torch.randn(n2, device=dev, generator=g) * 0.1
    y = torch.full((b, h, w, 256), float("nan"), device=dev, dtype=torch.bfloat16)
    y2 = torch.full((b, h, w, n2), float("nan"), device=dev, dtype=torch.bfloat16)
    op = ops.Igemm(ops.IGEMM_CONV, nhwc(x), ops.pack_conv_weight(wt, 256), sc, sh, y, batch=b, in_h=h, in_w=w, in_c=64, out_c=256,
                   ksize=1, stride=1, flags=ops.EPI_RELU | ops.EPI_RES, residual=nhwc(r), res_ld=256,
                   chain=(ops.pack_conv_weight(w2, n2), s2, t2, y2, True), name="chain")
    op()
    op()
    torch.cuda.synchronize()
    xb, wb, rb = x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), r.to(torch.bfloat16).float()
    ref = (F.conv2d(xb, wb) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1) + rb).relu()
    close(nchw(y), ref)
    ref2 = (F.conv2d(nchw(y), w2.to(torch.bfloat16).float()) * s2.view(1, -1, 1, 1) + t2.view(1, -1, 1, 1)).relu()
    close(nchw(y2), ref2)
    # and the unchained launch gives the same y bit for bit
    y0 = torch.empty_like(y)
    ops.Igemm(ops.IGEMM_CONV, nhwc(x), ops.pack_conv_weight(wt, 256), sc, sh, y0, batch=b, in_h=h, in_w=w, in_c=64, out_c=256,
              ksize=1, stride=1, flags=ops.EPI_RELU | ops.EPI_RES, residual=nhwc(r), res_ld=256, name="plain")()
    torch.cuda.synchronize()
    assert torch.equal(y0, y)


SEG2_CASES = [
    # (batch, H, W, C1, C2, Cout, stride2, chain_n or 0)
    (2, 24, 40, 64, 64, 256, 1, 0),      # layer1.0: conv3 over [t | x]
    (2, 24, 40, 64, 64, 256, 1, 64),     # ... with the next conv1 chained on (no residual on this path)
    (1, 20, 36, 128, 256, 512, 2, 0),    # layer2.0: projection input at stride 2, two n-blocks
    (2, 16, 16, 256, 512, 1024, 2, 0),   # layer3.0
    (1, 12, 20, 512, 1024, 2048, 2, 0),  # layer4.0: 24 k-blocks, CTA pair
    (16, 32, 32, 512, 1024, 2048, 2, 0),
]


@pytest.mark.parametrize("case", SEG2_CASES, ids=lambda c: "b%d_%dx%d_c%d+%d_o%d_s%d_ch%d" % c)
def test_second_k_segment(case):
    """b200_igemm_desc.x2: y = relu([x | x2 at stride2] . w^T + shift) -- a Bottleneck's projection shortcut inside conv3."""
    from ocr_vi_invoice_b200 import ops
    _setup()
    b, h, w, c1, c2, co, s2, n2 = case
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(c1 + c2 + h)
    x = torch.randn(b, c1, h, w, device=dev, generator=g)
    x2 = torch.randn(b, c2, h * s2, w * s2, device=dev, generator=g)
    w1 = torch.randn(co, c1, 1, 1, device=dev, generator=g) / (c1 + c2) ** 0.5
    w2 = torch.randn(co, c2, 1, 1, device=dev, generator=g) / (c1 + c2) ** 0.5
    sc, sh = torch.rand(co, device=dev, generator=g) + 0.5, torch.randn(co, device=dev, generator=g) * 0.1
    wp = torch.cat([ops.pack_conv_weight(w1, co), ops.pack_conv_weight(w2, co)], dim=1).contiguous()
    y = torch.full((b, h, w, co), float("nan"), device=dev, dtype=torch.bfloat16)
    chain, y2 = None, None
    if n2:
        wc = torch.randn(n2, co, 1, 1, device=dev, generator=g) / co ** 0.5
        sc2, sh2 = torch.rand(n2, device=dev, generator=g) + 0.5, torch.randn(n2, device=dev, generator=g) * 0.1
        y2 = torch.full((b, h, w, n2), float("nan"), device=dev, dtype=torch.bfloat16)
        chain = (ops.pack_conv_weight(wc, n2), sc2, sh2, y2, True)
    op = ops.Igemm(ops.IGEMM_CONV, nhwc(x), wp, sc, sh, y, batch=b, in_h=h, in_w=w, in_c=c1, out_c=co, ksize=1, stride=1,
                   flags=ops.EPI_RELU, x2=nhwc(x2), in_c2=c2, stride2=s2, chain=chain, name=str(case))
    op()
    op()
    torch.cuda.synchronize()
    bfr = lambda t: t.to(torch.bfloat16).float()
    ref = F.conv2d(bfr(x), bfr(w1)) + F.conv2d(bfr(x2), bfr(w2), stride=s2)
    ref = (ref * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)).relu()
    close(nchw(y), ref)
    if n2:
        ref2 = (F.conv2d(nchw(y), bfr(wc)) * sc2.view(1, -1, 1, 1) + sh2.view(1, -1, 1, 1)).relu()
        close(nchw(y2), ref2)


DCN_CASES = [
    # (batch, Hin, Win, C, Cout, stride, offset sigma)
    (1, 16, 16, 64, 64, 1, 0.0),
    (1, 16, 16, 64, 64, 1, 2.0),
    (2, 32, 24, 128, 128, 1, 2.0),
    (1, 40, 48, 128, 128, 2, 2.0),
    (2, 20, 20, 256, 256, 1, 3.0),
    (1, 32, 32, 256, 256, 2, 3.0),
    (1, 10, 14, 512, 512, 1, 4.0),
    (3, 48, 80, 128, 128, 1, 2.0),    # several tiles per CTA-less grid: ring phases wrap, ragged tile columns
    (16, 64, 64, 256, 256, 1, 3.0),   # 512 tiles on 148 SMs: 3-4 tiles per CTA, record double buffer recycled
]


@pytest.mark.parametrize("case", DCN_CASES, ids=lambda c: "b%d_%dx%d_c%d_o%d_s%d_sig%g" % c)
def test_dcn_igemm(case):
    import torchvision.ops as tvops
    from ocr_vi_invoice_b200 import ops
    _setup()
    b, h, w, ci, co, s, sig = case
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(sum(int(v * 7) for v in case if not isinstance(v, str)) & 0xffff)
    x = torch.randn(b, ci, h, w, device=dev, generator=g)
    wt = torch.randn(co, ci, 3, 3, device=dev, generator=g) / (9 * ci) ** 0.5
    oh, ow = h // s, w // s
    om = torch.zeros(b, oh, ow, 32, device=dev)
    om[..., :18] = torch.randn(b, oh, ow, 18, device=dev, generator=g) * sig
    om[..., 18:27] = torch.randn(b, oh, ow, 9, device=dev, generator=g)
    # a few exact-integer and boundary offsets (bilinear edge cases: h == -1, h == H, integer coords)
    om[:, 0, 0, 0:18] = -1.0
    om[:, -1, -1, 0:18] = 1.0
    om[:, oh // 2, ow // 2, 0:18] = 0.0
    scale = torch.rand(co, device=dev, generator=g) + 0.5
    shift = torch.randn(co, device=dev, generator=g) * 0.1
    xb, wb = x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float()
    offset = om[..., :18].permute(0, 3, 1, 2).contiguous()
    mask = torch.sigmoid(om[..., 18:27]).permute(0, 3, 1, 2).contiguous()
    ref = tvops.deform_conv2d(xb, offset, wb, None, stride=s, padding=1, mask=mask)
    ref = (ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).relu()
    wp = ops.pack_conv_weight(wt, ops.choose_n_pad(co))   # K = (chunk, tap, 64), same for both kernels
    wp_ref = wp
    n_pad = wp.shape[0]
    sc, sh = ops.pad_vec(scale, n_pad, 1.0), ops.pad_vec(shift, n_pad)
    y = torch.full((b, oh, ow, co), float("nan"), device=dev, dtype=torch.bfloat16)
    coords = torch.empty(b, oh, ow, 9, 4, device=dev, dtype=torch.int32)
    ops.dcn_coords(om, coords, b, h, w, ci, s)
    op = ops.Igemm(ops.IGEMM_DCN, nhwc(x), wp, sc, sh, y, batch=b, in_h=h, in_w=w, in_c=ci, out_c=co, ksize=3,
                   stride=s, flags=ops.EPI_RELU, offsets=coords, name=str(case))
    op()
    torch.cuda.synchronize()
    close(nchw(y), ref, rel_rms=3e-2)
    y2 = torch.zeros_like(y)
    ops.ref_dcn2d(nhwc(x), om, wp_ref, sc, sh, y2, b, h, w, ci, co, s, ops.EPI_RELU, co)
    torch.cuda.synchronize()
    close(nchw(y2), ref, rel_rms=3e-2)
    if co in (128, 256):
        # the tensor-memory kernel (A operand in TMEM, K permuted inside each chunk): the same blended bf16 A values, so
        # it differs from the first kernel only by the fp32 accumulation order over K
        y3 = torch.full_like(y, float("nan"))
        op3 = ops.Igemm(ops.IGEMM_DCN_TMEM, nhwc(x), ops.pack_dcn_weight(wt, n_pad), sc, sh, y3, batch=b, in_h=h, in_w=w,
                        in_c=ci, out_c=co, ksize=3, stride=s, flags=ops.EPI_RELU, offsets=coords, name="tmem " + str(case))
        op3()
        op3()
        torch.cuda.synchronize()
        close(nchw(y3), ref, rel_rms=3e-2)
        d3 = (nchw(y3) - nchw(y)).float()
        assert d3.pow(2).mean().sqrt().item() < 2e-3 * nchw(y).float().pow(2).mean().sqrt().item()
    # tensor-core path vs cross-check kernel: the producers blend the corners with packed bf16x2 FMAs (bf16 weights,
    # one rounding per FMA) where the cross-check kernel blends in fp32, so a few outputs differ by 1-2 bf16 ulps
    diff = (nchw(y) - nchw(y2)).float()
    rms = nchw(y2).float().pow(2).mean().sqrt().item()
    assert diff.pow(2).mean().sqrt().item() < 5e-3 * rms, "bf16x2 blend noise above 0.5% of the output rms"
    assert not (diff.abs() > 4e-2 * rms + 2.0 ** -6 * nchw(y2).float().abs()).any()  # blend noise + two bf16 ulps


@pytest.mark.parametrize("case", [(1, 24, 40, 128, 1), (2, 32, 48, 256, 2)], ids=lambda c: "b%d_%dx%d_c%d_s%d" % c)
def test_offset_conv_coords_epilogue(case):
    """The offset conv's fused sampling-record epilogue equals raw fp32 output -> b200_dcn_coords, bit for bit."""
    from ocr_vi_invoice_b200 import ops
    _setup()
    b, h, w, ci, s = case
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(21)
    x = torch.randn(b, ci, h, w, device=dev, generator=g)
    wt = torch.randn(27, ci, 3, 3, device=dev, generator=g) * (2.0 / (9 * ci) ** 0.5)
    bias = torch.randn(27, device=dev, generator=g)
    oh, ow = h // s, w // s
    wp = ops.pack_conv_weight(wt, 32)
    sc, sh = torch.ones(32, device=dev), ops.pad_vec(bias, 32)
    raw = torch.empty(b, oh, ow, 32, device=dev, dtype=torch.float32)
    ops.Igemm(ops.IGEMM_CONV, nhwc(x), wp, sc, sh, raw, batch=b, in_h=h, in_w=w, in_c=ci, out_c=32, ksize=3, stride=s,
              flags=ops.EPI_OUT_F32, name="offset-raw")()
    c_ref = torch.empty(b, oh, ow, 9, 4, device=dev, dtype=torch.int32)
    ops.dcn_coords(raw, c_ref, b, h, w, ci, s)
    c_fused = torch.full((b, oh, ow, 9, 4), -7, device=dev, dtype=torch.int32)
    ops.Igemm(ops.IGEMM_CONV, nhwc(x), wp, sc, sh, c_fused, batch=b, in_h=h, in_w=w, in_c=ci, out_c=32, ksize=3, stride=s,
              flags=ops.EPI_DCN_COORDS, out_ld=32, name="offset-coords")()
    torch.cuda.synchronize()
    assert torch.equal(c_fused, c_ref)
    # records are sane: the clamped 2x2 block lies inside the image, weights in [0, 1], each tap's sum <= sigmoid(mask) <= 1
    off = c_ref[..., 0]
    wg = c_ref[..., 1:3].contiguous().view(torch.bfloat16).float()
    assert off.min() >= 0 and off.max() <= ((h - 2) * w + (w - 2)) * ci
    assert wg.min() >= 0 and wg.sum(-1).max() <= 1.0 + 1e-2
