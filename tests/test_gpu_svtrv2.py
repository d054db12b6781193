"""GPU: SVTRv2 recogniser -- the svtr.cu kernels one by one against torch fp32, then the whole
drop-in model against the oracle (pinned to the reference's own module) on the golden crops.

Tolerances (bf16 storage, fp32 accumulation, stated per check): single kernels 1 bf16 ulp-ish (2^-7 relative to the row
scale); whole model: rms(error) / rms(stage) per stage tap, and agreement of the per-frame arg-max where the oracle's
margin between the two best classes exceeds the observed logit error."""
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
DEV = "cuda"


def _bf(t):
    return t.to(torch.bfloat16)


def _rel(a, b):
    return float((a.float() - b.float()).pow(2).mean().sqrt() / b.float().pow(2).mean().sqrt().clamp_min(1e-12))


def test_layernorm_gelu_decimate_logsoftmax():
    from ocr_vi_invoice_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(0)
    for d in (128, 256, 384, 1024):
        rows = 777
        x = _bf(torch.randn(rows, d, generator=g) * 3 + 1).to(DEV)
        gamma, beta = (torch.rand(d, generator=g) + 0.5).to(DEV), torch.randn(d, generator=g).to(DEV)
        y = torch.empty_like(x)
        ops.layernorm(x, y, gamma, beta, rows, d, 1e-5)
        ref = F.layer_norm(x.float(), (d,), gamma, beta, 1e-5)
        assert torch.equal(y, _bf(ref)) or (y.float() - ref).abs().max() <= 2 ** -7 * ref.abs().max()
        assert _rel(y, ref) < 3e-3
    x = _bf(torch.randn(4096 * 8, generator=g) * 3).to(DEV)
    r = _bf(torch.randn(4096 * 8, generator=g)).to(DEV)
    y = torch.empty_like(x)
    ops.gelu(x, y)
    assert _rel(y, F.gelu(x.float())) < 3e-3
    ops.gelu(x, y, r)
    assert _rel(y, F.gelu(x.float()) + r.float()) < 3e-3
    ops.gelu(x)  # in place
    assert _rel(x, y.float() - r.float()) < 1e-2
    t = _bf(torch.randn(3, 8, 64, 128, generator=g)).to(DEV)
    o = torch.empty(3, 4, 64, 128, dtype=torch.bfloat16, device=DEV)
    ops.decimate_rows(t, o, 3, 8, 64, 128)
    assert torch.equal(o, t[:, ::2])
    logits = torch.randn(5 * 64, 232, generator=g).to(DEV) * 4
    out = torch.empty(64, 5, 232, device=DEV)
    ops.log_softmax_tbc(logits, out, 5, 64, 232, 232)
    ref = F.log_softmax(logits.reshape(5, 64, 232).permute(1, 0, 2), dim=-1)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)


def test_stem_conv1():
    from ocr_vi_invoice_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 3, 32, 256, generator=g)
    w = torch.randn(64, 3, 3, 3, generator=g) * 0.3
    scale, shift = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.2
    y = torch.empty(3, 16, 128, 64, dtype=torch.bfloat16, device=DEV)
    ops.svtr_stem1(x.to(DEV), w.reshape(64, 27).contiguous().to(DEV), scale.to(DEV), shift.to(DEV), y)
    ref = F.gelu(F.conv2d(x, w, None, 2, 1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    assert _rel(y.cpu(), ref) < 3e-3


@pytest.mark.parametrize("n_seq,n,d", [(6, 64, 384), (3, 256, 256), (2, 128, 384), (1, 512, 128), (5, 40, 64)])
def test_mhsa(n_seq, n, d):
    from ocr_vi_invoice_b200 import ops
    g = torch.Generator().manual_seed(n)
    qkv = _bf(torch.randn(n_seq, n, 3 * d, generator=g) * 1.5).to(DEV)
    out = torch.empty(n_seq, n, d, dtype=torch.bfloat16, device=DEV)
    ops.mhsa(qkv, out, n_seq, n, d)
    heads = d // 32
    t = qkv.float().reshape(n_seq, n, 3, heads, 32).permute(2, 0, 3, 1, 4)
    ref = ((t[0] @ t[1].transpose(-2, -1)) * 32 ** -0.5).softmax(-1) @ t[2]
    ref = ref.transpose(1, 2).reshape(n_seq, n, d)
    assert _rel(out, ref) < 4e-3


def test_frm_select():
    from ocr_vi_invoice_b200 import ops
    g = torch.Generator().manual_seed(2)
    b, h, w, d = 3, 2, 64, 384
    q = torch.randn(d, generator=g).to(DEV)
    kv = _bf(torch.randn(b, h, w, 2 * d, generator=g)).to(DEV)
    out = torch.empty(b, w, d, dtype=torch.bfloat16, device=DEV)
    ops.frm_select(q, kv, out, b, h, w, d)
    heads = d // 32
    k = kv.float()[..., :d].reshape(b, h, w, heads, 32)
    v = kv.float()[..., d:].reshape(b, h, w, heads, 32)
    s = (k * q.reshape(heads, 32)).sum(-1) * 32 ** -0.5          # (b, h, w, heads)
    p = s.softmax(dim=1)
    ref = (p.unsqueeze(-1) * v).sum(1).reshape(b, w, d)
    assert _rel(out, ref) < 4e-3


def test_linear_as_token_image_gemm():
    """nn.Linear through the tcgen05 kernel with N not a multiple of the block (232 fp32 logits, 1152 / 384 bf16 columns)."""
    from ocr_vi_invoice_b200 import ops
    g = torch.Generator().manual_seed(3)
    rows = 6 * 64
    for din, dout, f32out in ((384, 232, True), (384, 1152, False), (1536, 384, False), (128, 512, False)):
        x = _bf(torch.randn(rows, din, generator=g)).to(DEV)
        wt = torch.randn(dout, din, generator=g) * din ** -0.5
        bias = torch.randn(dout, generator=g)
        w = ops.pack_conv_weight(wt.reshape(dout, din, 1, 1).to(DEV))
        s, t = torch.ones(w.shape[0], device=DEV), ops.pad_vec(bias.to(DEV), w.shape[0])
        res = _bf(torch.randn(rows, dout, generator=g)).to(DEV)
        y = torch.full((rows, dout), 7.0, dtype=torch.float32 if f32out else torch.bfloat16, device=DEV)
        flags = ops.EPI_OUT_F32 if f32out else ops.EPI_RES
        op = ops.Igemm(ops.IGEMM_CONV, x, w, s, t, y, batch=1, in_h=rows // 64, in_w=64, in_c=din, out_c=dout, ksize=1, stride=1,
                       flags=flags, residual=None if f32out else res, res_ld=0 if f32out else dout)
        op()
        ref = x.float() @ _bf(wt).float().t().to(DEV) + bias.to(DEV) + (0 if f32out else res.float())
        assert _rel(y, ref) < (1e-4 if f32out else 4e-3), (din, dout)


def test_gelu_epilogue_conv_and_linear():
    """B200_EPI_GELU: GELU(scale * acc + shift) [+ residual, added AFTER the GELU] in the tcgen05 kernel's epilogue, for a
    Linear (TMA-fed variant), a 3x3 stride-1 conv with N = 128 (share variant) and a stride-2 conv."""
    from ocr_vi_invoice_b200 import ops
    g = torch.Generator().manual_seed(4)
    # Linear 128 -> 512 + GELU
    rows, din, dout = 4 * 64, 128, 512
    x = _bf(torch.randn(rows, din, generator=g)).to(DEV)
    wt, bias = torch.randn(dout, din, generator=g) * din ** -0.5, torch.randn(dout, generator=g)
    w = ops.pack_conv_weight(wt.reshape(dout, din, 1, 1).to(DEV))
    y = torch.empty(rows, dout, dtype=torch.bfloat16, device=DEV)
    ops.Igemm(ops.IGEMM_CONV, x, w, torch.ones(w.shape[0], device=DEV), ops.pad_vec(bias.to(DEV), w.shape[0]), y, batch=1,
              in_h=rows // 64, in_w=64, in_c=din, out_c=dout, ksize=1, stride=1, flags=ops.EPI_GELU)()
    ref = F.gelu(x.float().cpu() @ _bf(wt).float().t() + bias)
    assert _rel(y.cpu(), ref) < 4e-3
    # 3x3 convs (+ residual after the GELU)
    for cin, cout, stride, hh, ww, res in ((128, 128, 1, 8, 64, True), (64, 128, 2, 16, 128, False), (256, 256, 1, 4, 64, True)):
        b = 2
        xi = _bf(torch.randn(b, hh, ww, cin, generator=g)).to(DEV)
        wt = torch.randn(cout, cin, 3, 3, generator=g) * (cin * 9) ** -0.5
        sc, sh = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.3
        w = ops.pack_conv_weight(wt.to(DEV))
        oh, ow = hh // stride, ww // stride
        r = _bf(torch.randn(b, oh, ow, cout, generator=g)).to(DEV)
        y = torch.empty(b, oh, ow, cout, dtype=torch.bfloat16, device=DEV)
        ops.Igemm(ops.IGEMM_CONV, xi, w, ops.pad_vec(sc.to(DEV), w.shape[0], 1.0), ops.pad_vec(sh.to(DEV), w.shape[0]), y, batch=b,
                  in_h=hh, in_w=ww, in_c=cin, out_c=cout, ksize=3, stride=stride, flags=ops.EPI_GELU | (ops.EPI_RES if res else 0),
                  residual=r if res else None, res_ld=cout if res else 0)()
        ref = F.conv2d(xi.float().cpu().permute(0, 3, 1, 2), _bf(wt).float(), None, stride, 1)
        ref = F.gelu(ref * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
        if res:
            ref = ref + r.float().cpu()
        assert _rel(y.cpu(), ref) < 4e-3, (cin, cout, stride)


def test_conv_stride_2_1():
    """PatchMerging's 3x3 conv with stride (2, 1) (two row-parity TMA views) vs torch, 128 -> 256 and 256 -> 384 channels."""
    from ocr_vi_invoice_b200 import ops
    g = torch.Generator().manual_seed(6)
    for cin, cout, hh, ww in ((128, 256, 8, 64), (256, 384, 4, 64), (64, 128, 16, 32)):
        b = 3
        xi = _bf(torch.randn(b, hh, ww, cin, generator=g)).to(DEV)
        wt = torch.randn(cout, cin, 3, 3, generator=g) * (cin * 9) ** -0.5
        sc, sh = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.3
        w = ops.pack_conv_weight(wt.to(DEV))
        y = torch.full((b, hh // 2, ww, cout), 7.0, dtype=torch.bfloat16, device=DEV)
        ops.Igemm(ops.IGEMM_CONV, xi, w, ops.pad_vec(sc.to(DEV), w.shape[0], 1.0), ops.pad_vec(sh.to(DEV), w.shape[0]), y, batch=b,
                  in_h=hh, in_w=ww, in_c=cin, out_c=cout, ksize=3, stride=2, stride_w=1)()
        ref = F.conv2d(xi.float().cpu().permute(0, 3, 1, 2), _bf(wt).float(), None, (2, 1), 1)
        ref = (ref * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
        assert _rel(y.cpu(), ref) < 4e-3, (cin, cout)


def _golden_model():
    from oracle import svtrv2_oracle as so
    from ocr_vi_invoice_b200.model.rec2.svtrv2 import SVTRv2
    g = np.load(GOLD / "svtrv2_small.npz")
    sd = so.make_state_dict("base", int(g["seed"]))
    model = SVTRv2(variant="base")
    model.load_state_dict(sd, strict=True)
    return g, sd, model.to(DEV).eval()


def test_model_against_oracle_and_reference_golden():
    from oracle import svtrv2_oracle as so
    g, sd, model = _golden_model()
    x = torch.from_numpy(g["x"])
    lp, taps = model.forward_with_taps(x.to(DEV))
    ref_taps = {}
    ref = so.svtrv2_forward(sd, x, "base", ref_taps)
    torch.testing.assert_close(ref, torch.from_numpy(g["log_probs"]), rtol=1e-4, atol=1e-4)  # the oracle is the reference
    errs = {k: _rel(taps[k].cpu(), ref_taps[k]) for k in ("stem", "stage0", "stage1", "stage2", "frm_h", "frm")}
    print("rel rms error per stage:", {k: round(v, 4) for k, v in errs.items()})
    assert errs["stem"] < 0.01 and errs["stage0"] < 0.02 and errs["stage1"] < 0.03 and errs["stage2"] < 0.04
    assert errs["frm_h"] < 0.05 and errs["frm"] < 0.05
    lp = lp.cpu()
    assert lp.shape == ref.shape and torch.isfinite(lp).all()
    torch.testing.assert_close(lp.exp().sum(-1), torch.ones(lp.shape[:2]), rtol=1e-4, atol=1e-4)
    err = (lp - ref).abs().max().item()
    print("max |log_prob error|:", err)
    assert err < 0.15
    top2 = ref.topk(2, dim=-1).values
    clear = (top2[..., 0] - top2[..., 1]) > 2 * err
    assert clear.float().mean() > 0.5
    assert torch.equal(lp.argmax(-1)[clear], ref.argmax(-1)[clear])


def test_model_other_crop_sizes():
    """(64, 128) and (32, 384) crops: other token grids (16 x 32 -> 4 x 32, 8 x 96 -> 2 x 96) against the oracle."""
    from oracle import svtrv2_oracle as so
    g, sd, model = _golden_model()
    gen = torch.Generator().manual_seed(7)
    for hh, ww in ((64, 128), (32, 384)):
        x = torch.randn(2, 3, hh, ww, generator=gen)
        lp = model(x.to(DEV)).cpu()
        ref = so.svtrv2_forward(sd, x, "base")
        assert lp.shape == ref.shape == (ww // 4, 2, 232)
        assert (lp - ref).abs().max().item() < 0.15, (hh, ww)
    with pytest.raises(ValueError, match="multiple of 32"):
        model(torch.zeros(1, 3, 40, 256, device=DEV))


@pytest.mark.parametrize("variant", ["tiny", "small"])
def test_other_variants_against_reference_golden(variant):
    """'tiny' (stem 32 -> 64 padded) and 'small' (stage width 96 -> 128, stem 48 -> 64 padded; the class default): zero-padded
    channels, LayerNorm over the real width; log-probabilities vs the reference module's (golden) and taps vs the oracle."""
    from oracle import svtrv2_oracle as so
    from ocr_vi_invoice_b200.model.rec2.svtrv2 import SVTRv2
    g = np.load(GOLD / "svtrv2_small.npz")
    sd = so.make_state_dict(variant, int(g["seed"]))
    model = SVTRv2(variant=variant)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    x = torch.from_numpy(g["x"])
    lp, taps = model.forward_with_taps(x.to(DEV))
    ref_taps = {}
    ref = so.svtrv2_forward(sd, x, variant, ref_taps)
    torch.testing.assert_close(ref, torch.from_numpy(g[f"log_probs_{variant}"]), rtol=1e-4, atol=1e-4)
    errs = {k: _rel(taps[k].cpu(), ref_taps[k]) for k in ("stem", "stage0", "stage1", "stage2", "frm")}
    print(variant, "rel rms error per stage:", {k: round(v, 4) for k, v in errs.items()})
    assert max(errs.values()) < 0.04
    assert (lp.cpu() - ref).abs().max().item() < 0.15


def test_model_batches_and_decode():
    """Odd batch (padding crop), batch invariance bit for bit, decode_greedy == decode_probs(forward), CPU refusal."""
    g, sd, model = _golden_model()
    x = torch.from_numpy(g["x"]).to(DEV)
    lp3 = model(x).clone()
    big = torch.cat([x, x.flip(0), x[:1]])  # 7 crops
    lp7 = model(big).clone()
    assert torch.equal(lp7[:, :3], lp3) and torch.equal(lp7[:, 6], lp3[:, 0]) and torch.equal(lp7[:, 3], lp3[:, 2])
    model.use_cuda_graph = True  # graph replay: same bits, twice
    assert torch.equal(model(x), lp3) and torch.equal(model(x), lp3)
    model.use_cuda_graph = False
    texts = model.decode_greedy(x)
    assert texts == model.decode_probs(lp3) and len(texts) == 3 and all(isinstance(t, str) for t in texts)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(x.cpu())
    with pytest.raises(RuntimeError, match="inference-only"):
        model(x, targets=torch.zeros(3, 4, dtype=torch.long))


def test_ocr_pipeline_page_to_text():
    """OCRPipeline (detect -> crop -> recognise -> decode on the device) == the same stages called one by one, and the
    rescaled boxes == the reference's in-place rescale (pipeline2.py:323-328) of the detector's boxes."""
    from oracle import recprep_oracle as ro
    from oracle import svtrv2_oracle as so
    from oracle.weights import make_state_dict
    from ocr_vi_invoice_b200.model.det.dbnet import DBNetPP
    from ocr_vi_invoice_b200.model.rec2.svtrv2 import SVTRv2
    from ocr_vi_invoice_b200.pipeline import OCRPipeline
    from ocr_vi_invoice_b200.postprocess import DBPostProcessor
    from ocr_vi_invoice_b200.preprocess import det_size, resize_image_for_det
    from ocr_vi_invoice_b200.synth import synthetic_batch_u8
    dev = torch.device(DEV, 0)
    det = DBNetPP(pretrained=False)
    det.load_state_dict(make_state_dict(seed=3, offset_px=0.5), strict=True)
    det = det.to(dev).eval()
    rec = SVTRv2(variant="base")
    rec.load_state_dict(so.make_state_dict("base", 5), strict=True)
    rec = rec.to(dev).eval()
    post = DBPostProcessor(thresh=0.3, box_thresh=0.5, unclip_ratio=1.6, device=dev)
    ocr = OCRPipeline(det, post, rec, dev, det_size=256)
    page_np = np.ascontiguousarray(synthetic_batch_u8(1, 300, 420, seed=4)[0])
    page = torch.from_numpy(page_np).to(dev)
    boxes, texts = ocr(page)
    assert len(boxes) == len(texts) and all(b.dtype == np.int32 and b.shape == (4, 2) for b in boxes)
    # stage by stage
    resized, (sh, sw) = resize_image_for_det(page, 256)
    det_boxes, _ = post(det(resized[None])["binary"][0])
    assert len(det_boxes) == len(boxes)
    for got, b in zip(boxes, det_boxes):
        assert np.array_equal(got, ro.rescale_box(b, sh, sw))
    if boxes:
        want, _ = ro.crops_for_recognition(page_np, det_boxes, sh, sw)
        n = len(boxes)
        x = torch.zeros(-(-n // 32) * 32, 3, 32, 256, device=dev)
        x[:n] = torch.from_numpy(want).to(dev)
        lp = rec(x)
        assert rec.decode_probs(lp)[:n] == texts or rec.decode_probs(rec(ocr.cropper._buf["out"][:x.shape[0]]))[:n] == texts


def test_ocr_pipeline_no_boxes():
    """A page on which nothing passes box_thresh: no crops, no recogniser launch, empty lists (pipeline2.py:332-334)."""
    from oracle import svtrv2_oracle as so
    from oracle.weights import make_state_dict
    from ocr_vi_invoice_b200.model.det.dbnet import DBNetPP
    from ocr_vi_invoice_b200.model.rec2.svtrv2 import SVTRv2
    from ocr_vi_invoice_b200.pipeline import OCRPipeline
    from ocr_vi_invoice_b200.postprocess import DBPostProcessor
    dev = torch.device(DEV, 0)
    det = DBNetPP(pretrained=False)
    det.load_state_dict(make_state_dict(seed=3, offset_px=0.5), strict=True)
    rec = SVTRv2(variant="base")
    rec.load_state_dict(so.make_state_dict("base", 5), strict=True)
    post = DBPostProcessor(thresh=0.3, box_thresh=1.5, device=dev)  # a mean probability can never reach 1.5
    ocr = OCRPipeline(det.to(dev).eval(), post, rec.to(dev).eval(), dev, det_size=128)
    boxes, texts = ocr(torch.full((90, 120, 3), 255, dtype=torch.uint8, device=dev))
    assert boxes == [] and texts == []


@pytest.mark.parametrize("variant", ["tiny", "small", "base"])
def test_reference_api_checks(variant):
    """Mirror of the reference's tests/test_model.py:252-321 (test_svtrv2_forward inference half, test_svtrv2_decode):
    default-initialised model, batch 4 of 32 x 256, (T, B, V) log-probabilities <= 0, decode_greedy gives B strings."""
    from ocr_vi_invoice_b200.model.rec2.svtrv2 import SVTRv2
    torch.manual_seed(0)
    model = SVTRv2(variant=variant, in_channels=3, dropout=0.0).to(DEV)
    model.eval()
    x = torch.randn(4, 3, 32, 256, device=DEV)
    with torch.no_grad():
        log_probs = model(x)
    assert log_probs.shape == (256 // 4, 4, model.tokenizer.num_classes)
    assert log_probs.max() <= 0, "Log probs must be <= 0"
    texts = model.decode_greedy(x[:2])
    assert isinstance(texts, list) and len(texts) == 2 and all(isinstance(t, str) for t in texts)
