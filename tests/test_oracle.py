"""CPU: the oracle restatement against golden vectors produced by the reference itself
(tests/golden/make_golden.py, run in the authoring container where /root/reference exists)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import dbnet_oracle as dbo
from oracle import postproc_oracle as ppo
from oracle.weights import make_state_dict, state_dict_keys, unique_param_table

GOLD = Path(__file__).parent / "golden"


def test_state_dict_key_set():
    keys = state_dict_keys()
    assert len(keys) == 754 and len(set(keys)) == 754
    assert len(unique_param_table()) == 410
    sd = make_state_dict(3)
    assert list(sd.keys()) == keys
    assert len({id(v) for v in sd.values()}) == 410            # 344 alias pairs share storage
    assert sd["backbone.layer1.0.weight"] is sd["backbone.model.conv1.weight"]
    assert sd["backbone.layer3.2.conv2.offset_mask_conv.bias"].shape == (27,)
    n_params = sum(v.numel() for k, v in {id(v): v for v in sd.values()}.items() if v.dtype.is_floating_point
                   and v.dim() > 0) - sum(v.numel() for k, v in sd.items()
                                          if k.startswith("backbone.model.") and "running_" in k) \
        - sum(v.numel() for k, v in sd.items() if not k.startswith("backbone.") and "running_" in k)
    assert n_params == 30106637                                 # reference parameter count


def test_model_oracle_matches_reference_golden():
    g = np.load(GOLD / "dbnet_small.npz")
    sd = make_state_dict(int(g["seed"]))
    taps = {}
    out = dbo.dbnet_forward(sd, torch.from_numpy(g["x"]), taps=taps)
    for k, v in out.items():
        ref = torch.from_numpy(g["out." + k])
        assert v.shape == ref.shape
        # fp32 CPU convolutions are not run-to-run identical (thread partitioning changes the summation order) and
        # thresh_binary amplifies logit differences by k = 50: 1e-4, observed 1e-6 .. 1e-5
        torch.testing.assert_close(v, ref, rtol=1e-4, atol=1e-4)
    tap_keys = [k[4:] for k in g.files if k.startswith("tap.")]
    assert len(tap_keys) == 24
    for k in tap_keys:
        ref = torch.from_numpy(g["tap." + k])
        torch.testing.assert_close(taps[k], ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))


def test_model_oracle_with_torchvision_op():
    """Same forward, but through the installed third-party op the reference calls (dcn.py:48)."""
    pytest.importorskip("torchvision")
    g = np.load(GOLD / "dbnet_small.npz")
    sd = make_state_dict(int(g["seed"]))
    out = dbo.dbnet_forward(sd, torch.from_numpy(g["x"]), deform=dbo.torchvision_deform)
    torch.testing.assert_close(out["bin_logits"], torch.from_numpy(g["out.bin_logits"]), rtol=1e-4, atol=1e-4)


def test_deform_conv_restatement_matches_torchvision_golden():
    d = np.load(GOLD / "dcn_torchvision.npz")
    for i in range(3):
        t = lambda k: torch.from_numpy(d[f"{i}.{k}"])
        y = dbo.deform_conv2d_ref(t("x"), t("offset"), t("weight"), t("mask"), stride=int(d[f"{i}.stride"]))
        torch.testing.assert_close(y, t("y"), rtol=1e-5, atol=1e-5)


def test_deform_conv_float64_property():
    """Zero offsets and unit mask reduce the op to a plain convolution."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, 7, 9, generator=g, dtype=torch.float64)
    w = torch.randn(5, 4, 3, 3, generator=g, dtype=torch.float64)
    for s in (1, 2):
        ho, wo = (7 + 2 - 3) // s + 1, (9 + 2 - 3) // s + 1
        y = dbo.deform_conv2d_ref(x, torch.zeros(1, 18, ho, wo, dtype=torch.float64), w,
                                  torch.ones(1, 9, ho, wo, dtype=torch.float64), stride=s)
        torch.testing.assert_close(y, torch.nn.functional.conv2d(x, w, stride=s, padding=1), rtol=1e-12, atol=1e-12)


# ----------------------------------------------------------------------------- post-processing oracle
def test_labels_scipy_vs_numpy_and_canonical_order():
    for seed in range(4):
        prob = ppo.render_prob_map(96, 128, seed=seed, n_boxes=10)
        fg = ppo.foreground(prob, 0.3)
        a, na = ppo.label_components(fg)
        b, nb = ppo.label_components_numpy(fg)
        assert na == nb and np.array_equal(a, b)
        firsts = [np.flatnonzero(a.ravel() == l)[0] for l in range(1, na + 1)]
        assert firsts == sorted(firsts)


def test_postproc_edge_cases():
    z = np.zeros((32, 64), np.float32)
    lab, n, recs = ppo.postprocess(z)
    assert n == 0 and len(recs) == 0 and not lab.any()
    full = np.full((32, 64), 0.9, np.float32)
    lab, n, recs = ppo.postprocess(full, unclip_ratio=1.5)
    assert n == 1 and len(recs) == 1 and recs[0]["npix"] == 32 * 64
    d = (63 * 31 * 1.5) / (2 * (63 + 31))
    assert recs[0]["x0"] == -int(d + 0.5) and recs[0]["x1"] == int(63 + d + 0.5)
    # diagonal touch is one component (8-connectivity); a one-row run is dropped (zero-area contour)
    m = np.zeros((16, 16), np.float32)
    m[2:6, 2:6] = 0.9
    m[6:10, 6:10] = 0.9
    m[12, 1:15] = 0.9
    lab, n, recs = ppo.postprocess(m, min_area=1)
    assert n == 2 and len(recs) == 1 and (recs[0]["bx0"], recs[0]["by0"], recs[0]["bx1"], recs[0]["by1"]) == (2, 2, 9, 9)
    # max_candidates caps the examined components (test.py:66-67)
    dots = np.zeros((64, 64), np.float32)
    dots[::4, ::4] = 0.0
    for y in range(0, 64, 8):
        for x in range(0, 64, 8):
            dots[y:y + 5, x:x + 5] = 0.95
    lab, n, recs = ppo.postprocess(dots, max_candidates=10, min_area=4)
    assert n == 64 and len(recs) == 10 and recs[-1]["label"] == 10
    # box_thresh: exact integer comparison
    faint = np.zeros((16, 32), np.float32)
    faint[4:10, 4:20] = 0.5
    assert len(ppo.postprocess(faint, box_thresh=0.5, min_area=1)[2]) == 1
    assert len(ppo.postprocess(faint, box_thresh=0.5000001, min_area=1)[2]) == 0


def test_boxes_agree_with_cv2_half_of_reference():
    """Reported agreement (not bit-exact): bounding rectangles vs the reference's cv2 contour path."""
    p = np.load(GOLD / "postproc_cv2.npz")
    rates = []
    for i in range(3):
        _, _, recs = ppo.postprocess(p[f"{i}.prob"], 0.3, 0.5, 1000, 1.6, 10)
        ours = [(r["x0"], r["y0"], r["x1"], r["y1"]) for r in recs]
        theirs = [tuple(int(v) for v in r) for r in p[f"{i}.cv2_rects"]]
        rates.append(ppo.match_rate(ours, theirs, 0.8))
        boxes, scores = ppo.records_to_boxes(recs)
        assert all(b.shape == (4, 2) and b.dtype == np.int64 and b.flags.writeable for b in boxes)
        assert all(0.5 <= s <= 1.0 for s in scores)
    assert min(rates) >= 0.9, rates


# ----------------------------------------------------------------------------- reference's own post-processor
def _reference_goldens():
    g = np.load(GOLD / "postproc_reference.npz")
    for i in range(int(g["n_maps"])):
        th, bt, ratio, min_area, clean = g[f"params_{i}"]
        rects = [(int(x), int(y), int(x + w - 1), int(y + h - 1)) for x, y, w, h in g[f"rects_{i}"]]
        yield g[f"prob_{i}"], float(th), float(bt), float(ratio), int(min_area), bool(clean), rects, g[f"scores_{i}"]


def test_clipper_offset_known_answers():
    """ClipperOffset restatement (oracle/clipper_offset.py) on cases with closed-form answers."""
    from oracle.clipper_offset import ET_CLOSEDPOLYGON, JT_ROUND, Polygon, PyclipperOffset
    rect = [[10, 20], [110, 20], [110, 50], [10, 50]]
    poly = Polygon(rect)
    assert poly.area == 3000.0 and poly.length == 260.0
    for order in (rect, rect[::-1]):                      # either orientation grows outward (FixOrientations)
        off = PyclipperOffset()
        off.AddPath(order, JT_ROUND, ET_CLOSEDPOLYGON)
        (path,) = off.Execute(7.3)
        p = np.array(path)
        assert (p[:, 0].min(), p[:, 1].min(), p[:, 0].max(), p[:, 1].max()) == (3, 13, 117, 57)  # Round(10 - 7.3) = 3 ...
        # every offset vertex is within delta (+ rounding) of the rectangle and at least delta - arc tolerance away
        dx = np.maximum(np.maximum(10 - p[:, 0], p[:, 0] - 110), 0)
        dy = np.maximum(np.maximum(20 - p[:, 1], p[:, 1] - 50), 0)
        dist = np.hypot(dx, dy)
        assert dist.max() <= 7.3 + 0.71 and dist.min() >= 7.3 - 0.25 - 0.71
    off = PyclipperOffset()
    off.AddPath(rect + [rect[0]], JT_ROUND, ET_CLOSEDPOLYGON)  # duplicated closing vertex is stripped
    assert len(off.Execute(7.3)[0]) == len(path)
    # a concave (L-shaped) polygon: the reflex vertex emits the 3-point notch Clipper's union later removes
    ell = [[0, 0], [40, 0], [40, 20], [20, 20], [20, 40], [0, 40]]
    off = PyclipperOffset()
    off.AddPath(ell, JT_ROUND, ET_CLOSEDPOLYGON)
    q = np.array(off.Execute(5.0)[0])
    assert [20, 20] in q.tolist() and (q[:, 0].min(), q[:, 1].min(), q[:, 0].max(), q[:, 1].max()) == (-5, -5, 45, 45)


def test_cv2_restatement_equals_reference_code():
    """oracle.postproc_oracle.cv2_reference_path == bounding rectangles of the boxes the reference's own
    DBPostProcessor.__call__ returned (tests/golden/make_postproc_reference.py), in the reference's order."""
    for prob, th, bt, ratio, min_area, _clean, rects, scores in _reference_goldens():
        got, got_scores = ppo.cv2_reference_path(prob, th, bt, 1000, ratio, min_area)
        assert got == rects
        assert np.allclose(got_scores, scores, rtol=0, atol=1e-12)


def test_contract_boxes_equal_reference_boxes_on_rectangular_components():
    """For axis-aligned rectangular components the contract (component box + closed-form un-clip) and the reference
    (contour -> approxPolyDP -> Clipper round offset -> cv2.boundingRect in the consumer) give identical rectangles
    and scores; on noisy / nested / diagonally-touching components they agree for > 90 % at IoU 0.8."""
    n_clean = 0
    for prob, th, bt, ratio, min_area, clean, rects, scores in _reference_goldens():
        _, _, recs = ppo.postprocess(prob, th, bt, 1000, ratio, min_area)
        ours = [(int(r["x0"]), int(r["y0"]), int(r["x1"]), int(r["y1"])) for r in recs]
        _, our_scores = ppo.records_to_boxes(recs)
        if clean:
            n_clean += 1
            assert sorted(ours) == sorted(rects)
            assert np.allclose(sorted(our_scores), sorted(scores), rtol=0, atol=2.0 ** -23)
            # the reference lists contours bottom-up (reverse raster order); ours is raster order of the first pixel
            assert ours == [r for r in sorted(rects, key=lambda r: (min(r[1], r[3]), r[0]))] or sorted(ours) == sorted(rects)
        else:
            assert ppo.match_rate(ours, rects, 0.8) >= 0.9
            assert sum(1 for r in rects if r in ours) >= 0.85 * len(rects)
    assert n_clean >= 2


# ----------------------------------------------------------------------------- pre-processing
def test_resize_restatement_matches_cv2_golden():
    from oracle import preprocess_oracle as pre
    g = np.load(GOLD / "resize_cv2.npz")
    for i, (h, w, dh, dw) in enumerate(g["cases"]):
        got = pre.resize_u8_linear(g[f"src_{i}"], int(dw), int(dh))
        assert np.array_equal(got, g[f"dst_{i}"]), f"case {i}: {h}x{w} -> {dh}x{dw}"
    try:  # live check against the installed cv2 on more shapes (the golden file is what pins it)
        import cv2
    except ImportError:
        return
    rng = np.random.default_rng(1)
    for h, w, dh, dw in [(333, 517, 640, 992), (480, 640, 1024, 1024), (1000, 700, 640, 448), (31, 33, 256, 256)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(pre.resize_u8_linear(img, dw, dh), cv2.resize(img, (dw, dh)))


def test_det_size_matches_reference_formula():
    from oracle import preprocess_oracle as pre
    from ocr_vi_invoice_b200 import preprocess
    for h, w, size in [(1400, 1000, 640), (333, 517, 640), (1000, 1000, 1024), (280, 400, 1280), (997, 1403, 1024)]:
        assert pre.det_size(h, w, size) == preprocess.det_size(h, w, size)
        (nh, nw), (sh, sw) = pre.det_size(h, w, size)
        assert nh % 32 == 0 and nw % 32 == 0 and max(nh, nw) in (size - 16, size, size + 16) or max(nh, nw) % 32 == 0
        assert sh == nh / h and sw == nw / w


# ------------------------------------------------------------------ detection -> recognition hand-over
def _recprep_golden():
    return np.load(GOLD / "recprep_reference.npz")


def test_recprep_restatement_matches_reference_code():
    """oracle/recprep_oracle.py vs the reference's own crop_image / preprocess_for_recognition / rescale loop
    (tests/golden/make_recprep_golden.py): rectangles, statuses and every float of the tensors, bit for bit."""
    from oracle import recprep_oracle as ro
    g = _recprep_golden()
    table = g["norm_table"]
    seen = set()
    for name in "abc":
        page, boxes = g[f"{name}_page"], g[f"{name}_boxes"]
        sh, sw = (float(v) for v in g[f"{name}_scales"])
        out, rects = ro.crops_for_recognition(page, boxes, sh, sw)
        assert np.array_equal(rects[:, :4], g[f"{name}_rects"])
        assert np.array_equal(rects[:, 4], g[f"{name}_status"])
        idx = g[f"{name}_tensor_index"]
        want = np.stack([table[c][idx[:, c]] for c in range(3)], axis=1)
        want[g[f"{name}_status"] != 0] = 0.0
        assert out.dtype == np.float32 and np.array_equal(out, want)
        seen |= set(rects[:, 4].tolist())
    assert seen == {ro.STATUS_OK, ro.STATUS_EMPTY, ro.STATUS_DEGENERATE}


def test_recprep_normalisation_table_is_the_reference_arithmetic():
    from oracle import recprep_oracle as ro
    g = _recprep_golden()
    ramp = np.repeat(np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 3, axis=2), 32, axis=0)
    t, st = ro.preprocess_for_recognition(ramp)
    assert st == ro.STATUS_OK and np.array_equal(t[:, 0, :], g["norm_table"])


def test_recprep_quirks():
    from oracle import recprep_oracle as ro
    # truncation toward zero, not floor, for negative coordinates (pipeline2.py:325-326)
    rb = ro.rescale_box(np.array([[-9, -9], [10, -9], [10, 10], [-9, 10]]), 0.64, 0.64)
    assert rb.min() == -14 and rb.dtype == np.int32
    # a negative corner shifts the crop, it does not shrink it (test.py:126-129)
    assert ro.crop_rect(100, 100, np.array([[-5, 3], [20, 3], [20, 9], [-5, 9]])) == (0, 3, 26, 7)
    # right / bottom borders cut
    assert ro.crop_rect(100, 100, np.array([[90, 95], [120, 95], [120, 130], [90, 130]])) == (90, 95, 10, 5)
    img = np.zeros((100, 100, 3), np.uint8)
    assert ro.crop_image(img, np.array([[120, 5], [130, 5], [130, 9], [120, 9]])).size == 0
    assert ro.rec_width(400, 2) == 0 and ro.rec_width(32, 256) == 256 and ro.rec_width(33, 300) == 290


# ------------------------------------------------------------------ greedy CTC decode (tail)
def test_ctc_restatement_matches_reference_code():
    """oracle/ctc_oracle.py vs SVTRv2.decode_probs of the reference (tests/golden/make_ctc_golden.py): ids and strings."""
    from oracle import ctc_oracle as co
    g = np.load(GOLD / "ctc_reference.npz")
    charset = "".join(g["charset"].tolist())
    table = co.make_id_to_token(charset)
    assert len(table) == int(g["num_classes"])
    for name in "abc":
        lp = g[f"{name}_log_probs"]
        ids = co.greedy_ids(lp)
        assert [len(i) for i in ids] == g[f"{name}_lengths"].tolist()
        for b, i in enumerate(ids):
            assert i == g[f"{name}_ids"][b, :len(i)].tolist()
        assert co.decode_ids(ids, table) == g[f"{name}_texts"].tolist()
    assert g["a_lengths"][0] == 0  # the all-blank sequence


# ------------------------------------------------------------------ SVTRv2 recogniser
def test_svtrv2_restatement_matches_reference_module():
    """oracle/svtrv2_oracle.py vs the reference's own SVTRv2('base') (tests/golden/make_svtrv2_golden.py): key set,
    log-probabilities, stage taps, decoded strings."""
    from oracle import ctc_oracle as co
    from oracle import svtrv2_oracle as so
    g = np.load(GOLD / "svtrv2_small.npz")
    shapes = so.param_shapes("base")
    assert len(shapes) == 318
    assert sum(int(np.prod(s)) for k, s in shapes.items()
               if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))) == 22630864
    sd = so.make_state_dict("base", int(g["seed"]))
    taps = {}
    lp = so.svtrv2_forward(sd, torch.from_numpy(g["x"]), "base", taps)
    torch.testing.assert_close(lp, torch.from_numpy(g["log_probs"]), rtol=1e-4, atol=1e-4)
    st = int(g["tap_stride"])
    for k in ("stem", "stage0", "stage1", "stage2"):
        ref = torch.from_numpy(g["tap." + k])
        torch.testing.assert_close(taps[k][:, ::st], ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))
    for k in ("frm", "logits"):
        ref = torch.from_numpy(g["tap." + k])
        torch.testing.assert_close(taps[k], ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))
    from ocr_vi_invoice_b200.model.rec2 import VOCAB
    assert co.decode_probs(lp.numpy(), VOCAB) == g["texts"].tolist()
    for variant in ("tiny", "small"):
        lpv = so.svtrv2_forward(so.make_state_dict(variant, int(g["seed"])), torch.from_numpy(g["x"]), variant)
        torch.testing.assert_close(lpv, torch.from_numpy(g[f"log_probs_{variant}"]), rtol=1e-4, atol=1e-4)
