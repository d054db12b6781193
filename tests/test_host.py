"""CPU: host-side mirror of the reference interface (constructor, state_dict, weight re-layout, error behaviour).
Mirrors tests/test_model.py:57-183 of the reference where it does not need a device."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from ocr_vi_invoice_b200 import ops
from ocr_vi_invoice_b200.model.det.dbnet import DBNetPP
from oracle.weights import make_state_dict, state_dict_keys


@pytest.fixture(scope="module")
def model():
    return DBNetPP(pretrained=False)


def test_parameter_counts(model):
    count = lambda m: sum(p.numel() for p in m.parameters() if p.requires_grad)
    total = count(model)
    assert count(model.backbone) + count(model.neck) + count(model.head) == total  # test_model.py:96
    assert total == 30106637
    assert model.backbone.out_channels == [256, 512, 1024, 2048]                    # test_model.py:162-169


def test_state_dict_is_reference_compatible(model):
    assert list(model.state_dict().keys()) == state_dict_keys()
    sd = make_state_dict(5)
    assert model.load_state_dict(sd, strict=True).missing_keys == []
    prefixed = {"module." + k: v for k, v in sd.items()}                             # pipeline2.py:49-50
    assert model.load_state_dict({k.replace("module.", "", 1): v for k, v in prefixed.items()}).unexpected_keys == []
    torch.testing.assert_close(model.backbone.layer2[1].conv2.offset_mask_conv.bias,
                               sd["backbone.model.layer2.1.conv2.offset_mask_conv.bias"])


def test_offset_conv_zero_init():
    m = DBNetPP(pretrained=False)
    om = m.backbone.model.layer3[0].conv2.offset_mask_conv                          # dcn.py:28-29
    assert om.weight.abs().sum() == 0 and om.bias.abs().sum() == 0 and om.stride == (2, 2)


def test_rejections(model):
    with pytest.raises(NotImplementedError):
        DBNetPP(backbone="resnet18", pretrained=False)
    with pytest.raises(NotImplementedError):
        DBNetPP(backbone="resnet101", pretrained=False)                              # backbone.py:21 raises too
    with pytest.warns(RuntimeWarning, match="does not download"):
        m = DBNetPP()                                                                # pretrained=True default: val.py:147
    assert len(m.state_dict()) == 754                                                # constructs, then load_state_dict
    model.eval()
    with pytest.raises(RuntimeError, match="CUDA only"):
        model(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="no eager forward"):
        model.backbone(torch.zeros(1, 3, 64, 64))


def test_dcn_false_constructs():
    m = DBNetPP(pretrained=False, dcn=False)
    assert isinstance(m.backbone.model.layer2[0].conv2, torch.nn.Conv2d)
    assert len(m.state_dict()) == 754 - 2 * 2 * 13


def test_pack_conv_weight_is_chunk_then_tap_major():
    w = torch.randn(5, 128, 3, 3)
    p = ops.pack_conv_weight(w, 32).float()
    assert p.shape == (32, 1152) and p[5:].abs().sum() == 0
    # channel 71 = chunk 1, lane 7; tap (1, 1) = 4  ->  K index (1 * 9 + 4) * 64 + 7
    torch.testing.assert_close(p[2, (1 * 9 + 4) * 64 + 7], w[2, 71, 1, 1].to(torch.bfloat16).float())


def test_projection_shortcut_folds_into_conv3(model):
    """engine.PreparedWeights: [t | x] . [s3 W3 | sd Wd]^T + (t3 + td) == bn3(conv3(t)) + bn_d(conv_d(x)) (resnet.py:155-161)."""
    from ocr_vi_invoice_b200.model.det.engine import PreparedWeights
    sd = make_state_dict(11)
    model.load_state_dict(sd, strict=True)
    model.eval()
    W = PreparedWeights(model, torch.device("cpu"))
    for li, stride in ((1, 1), (2, 2)):
        blk = model.backbone.model.__getattr__(f"layer{li}")[0]
        wf, scale, shift = W.P[f"layer{li}.0.conv3d"]
        c1, c2 = blk.conv3.weight.shape[1], blk.downsample[0].weight.shape[1]
        assert wf.shape == (blk.conv3.weight.shape[0], c1 + c2) and (scale == 1).all()
        t = torch.randn(2, c1, 6, 5)
        x = torch.randn(2, c2, 6 * stride, 5 * stride)
        with torch.no_grad():
            ref = blk.bn3(blk.conv3(t)) + blk.downsample(x)
        k = torch.cat([t, x[:, :, ::stride, ::stride]], dim=1)              # the GEMM's K axis: conv3 input, then the block input
        got = torch.einsum("nkhw,ok->nohw", k, wf.float()) + shift.view(1, -1, 1, 1)
        err = (got - ref).abs().max().item()
        assert err < 0.03 * ref.abs().max().item() + 1e-3, (li, err)         # bf16 weights vs fp32 modules


def test_pack_stem_weight_equals_conv7x7():
    """The space-to-depth re-layout reproduces conv 7x7 s2 p3 exactly (fp32 emulation of pack + GEMM)."""
    g = torch.Generator().manual_seed(0)
    img = torch.randn(1, 3, 16, 24, generator=g).to(torch.bfloat16).float()
    w = torch.randn(64, 3, 7, 7, generator=g).to(torch.bfloat16).float()
    wp = ops.pack_stem_weight(w, 64).float()                       # [64][ty][j][16]
    H, W = 16, 24
    packed = torch.zeros(H // 2 + 3, W // 2 + 3, 16)
    blocks = img[0].reshape(3, H // 2, 2, W // 2, 2).permute(1, 3, 2, 4, 0).reshape(H // 2, W // 2, 12)
    packed[2:2 + H // 2, 2:2 + W // 2, :12] = blocks
    out = torch.zeros(64, H // 2, W // 2)
    for oy in range(H // 2):
        for ox in range(W // 2):
            a = packed[oy:oy + 4, ox:ox + 4, :].reshape(256)
            out[:, oy, ox] = wp @ a
    ref = F.conv2d(img, w, stride=2, padding=3)[0]
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)


def test_pack_convT_weight():
    w = torch.randn(64, 64, 2, 2)
    p = ops.pack_convT_weight(w, 256).float()
    torch.testing.assert_close(p[(1 * 2 + 0) * 64 + 9, 17], w[17, 9, 1, 0].to(torch.bfloat16).float())


def test_postprocessor_attributes():
    from ocr_vi_invoice_b200.postprocess import DBPostProcessor
    pp = DBPostProcessor(thresh=0.3, box_thresh=0.5, unclip_ratio=1.6)
    pp.min_area = 10                                                                  # pipeline2.py:254-259
    assert (pp.thresh, pp.box_thresh, pp.max_candidates, pp.unclip_ratio, pp.min_size) == (0.3, 0.5, 1000, 1.6, 3)
    with pytest.raises(RuntimeError):
        pp.run_device(torch.zeros(1, 8, 8))


def test_rec2_vocab_and_tokenizer_match_reference():
    """VOCAB (built from its description) == the charset of the reference's tokenizer; Tokenizer.decode of the golden id
    lists == the strings the reference's decode_probs returned (tests/golden/ctc_reference.npz)."""
    from pathlib import Path
    import numpy as np
    from ocr_vi_invoice_b200.model.rec2 import VOCAB, Tokenizer
    g = np.load(Path(__file__).parent / "golden" / "ctc_reference.npz")
    assert VOCAB == "".join(g["charset"].tolist())
    tok = Tokenizer(VOCAB)
    assert tok.num_classes == int(g["num_classes"]) and tok.blank_id == 0 and tok.pad_id == 1
    for name in "abc":
        ids = [row[:n].tolist() for row, n in zip(g[f"{name}_ids"], g[f"{name}_lengths"])]
        assert tok.decode(ids) == g[f"{name}_texts"].tolist()
    enc = tok.encode(["ab", "đồng ₫", ""])
    assert enc.shape == (3, 6) and enc[2].tolist() == [1] * 6 and tok.decode(enc) == ["ab", "đồng ₫", ""]
    assert tok.encode([]).shape == (0,) or tok.encode([]).numel() == 0


def test_rec2_decode_refuses_cpu_tensors():
    import pytest
    import torch
    from ocr_vi_invoice_b200.model.rec2 import VOCAB, Tokenizer, decode_probs
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        decode_probs(torch.zeros(4, 2, 232), Tokenizer(VOCAB))


def test_svtrv2_dropin_state_dict_matches_reference_key_set():
    """The drop-in SVTRv2 holds exactly the reference's 318 state_dict entries (names, order, shapes: the oracle's table is
    checked against the real module by tests/golden/make_svtrv2_golden.py) and loads them strictly; unsupported variants
    are refused instead of running unaccelerated."""
    import pytest
    from oracle import svtrv2_oracle as so
    from ocr_vi_invoice_b200.model.rec2.svtrv2 import SVTRv2
    model = SVTRv2(variant="base")
    sd = model.state_dict()
    shapes = so.param_shapes("base", model.tokenizer.num_classes)
    assert list(sd.keys()) == list(shapes.keys())
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)
    model.load_state_dict(so.make_state_dict("base", 1), strict=True)
    assert model.blank_id == 0 and model.tokenizer.num_classes == 232
    for variant in ("tiny", "small"):  # the class default is 'small' (reference svtrv2.py:421)
        mv = SVTRv2(variant=variant) if variant == "tiny" else SVTRv2()
        assert list(mv.state_dict().keys()) == list(so.param_shapes(variant).keys())
    with pytest.raises(NotImplementedError):
        SVTRv2(in_channels=1)
    import torch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.eval()(torch.zeros(1, 3, 32, 256))


def test_svtrv2_parameter_counts_like_reference_test():
    """Mirror of the reference's tests/test_model.py:190-249 (test_svtrv2_parameters): per-component trainable parameter
    counts of the three variants (numbers produced by the reference's own classes)."""
    from ocr_vi_invoice_b200.model.rec2.svtrv2 import SVTRv2
    count = lambda mod: sum(p.numel() for p in mod.parameters() if p.requires_grad)
    want = {
        "tiny": (6736656, 19584, [211776, 1214976, 2369280], [74112, 295680], 512, 1580288, 910824, 59624),
        "small": (10893744, 43200, [391392, 2559744, 4738560], [166464, 443136], 512, 1580288, 910824, 59624),
        "base": (22630864, 76032, [620160, 4510208, 10646784], [295680, 885888], 768, 3550080, 1955944, 89320),
    }
    for v, (total, stem, stages, merges, bnorm, frm, sgm, head) in want.items():
        m = SVTRv2(variant=v, in_channels=3, dropout=0.0)
        assert count(m) == total and count(m.stem) == stem
        assert [count(s) for s in m.stages] == stages and [count(x) for x in m.merges] == merges
        assert count(m.backbone_norm) == bnorm and count(m.frm) == frm and count(m.sgm) == sgm and count(m.head) == head
