"""GPU parity of the post-processing (threshold -> CCL -> boxes): bit-exact against oracle/postproc_oracle.py."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def _run(prob_np, **kw):
    from ocr_vi_invoice_b200.postprocess import DBPostProcessor
    pp = DBPostProcessor(thresh=kw.get("thresh", 0.3), box_thresh=kw.get("box_thresh", 0.6),
                         max_candidates=kw.get("max_candidates", 1000), unclip_ratio=kw.get("unclip_ratio", 1.5))
    pp.min_area = kw.get("min_area", 10)
    prob = torch.from_numpy(prob_np).cuda()
    if prob.dim() == 2:
        prob = prob[None]
    buf = pp.run_device(prob, want_labels=True)
    recs, ncomp = pp.fetch(buf)
    return buf["labels"].cpu().numpy(), ncomp, recs, pp


def _assert_same(prob_np, **kw):
    from oracle import postproc_oracle as ppo
    labels, ncomp, recs, _ = _run(prob_np, **kw)
    maps = prob_np if prob_np.ndim == 3 else prob_np[None]
    for i, prob in enumerate(maps):
        ref_lab, ref_n, ref_recs = ppo.postprocess(prob, kw.get("thresh", 0.3), kw.get("box_thresh", 0.6),
                                                   kw.get("max_candidates", 1000), kw.get("unclip_ratio", 1.5),
                                                   kw.get("min_area", 10))
        assert ncomp[i] == ref_n
        assert np.array_equal(labels[i], ref_lab), "label map differs"
        assert len(recs[i]) == len(ref_recs)
        for name in ref_recs.dtype.names:
            assert np.array_equal(recs[i][name], ref_recs[name]), f"box field {name} differs"
    return recs


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_rendered_maps_bit_exact(seed):
    from oracle import postproc_oracle as ppo
    prob = np.stack([ppo.render_prob_map(256, 320, seed=seed * 10 + i, n_boxes=20 + 5 * i) for i in range(3)])
    recs = _assert_same(prob, thresh=0.3, box_thresh=0.5, unclip_ratio=1.6, min_area=10)
    assert sum(len(r) for r in recs) > 30


def test_fractional_min_area_like_reference():
    """pipeline2.py passes --det_min_area as a float: 10.5 must drop a box of area 10 (test.py:81 `area < min_area`)."""
    from oracle import postproc_oracle as ppo
    prob = np.zeros((64, 96), np.float32)
    prob[10:13, 10:16] = 0.9    # 6 x 3 pixels: (w-1)(h-1) = 10
    prob[30:34, 30:40] = 0.9    # 10 x 4 pixels: area 27
    r10 = _assert_same(prob, min_area=10, box_thresh=0.5)
    r105 = _assert_same(prob, min_area=10.5, box_thresh=0.5)
    assert len(r10[0]) == 2 and len(r105[0]) == 1


def test_edge_cases_bit_exact():
    rng = np.random.default_rng(0)
    _assert_same(np.zeros((64, 96), np.float32))                                   # empty map
    _assert_same(np.full((64, 96), 0.9, np.float32))                               # one giant component
    _assert_same(rng.random((128, 160), dtype=np.float32), thresh=0.5, min_area=1)  # noise: thousands of specks
    _assert_same(rng.random((70, 45), dtype=np.float32), thresh=0.4, min_area=1)   # width not a multiple of 32
    spiral = np.zeros((65, 65), np.float32)                                        # long thin snake, many merges
    for k in range(0, 32, 2):
        spiral[k, k:65 - k] = 0.9; spiral[64 - k, k:65 - k] = 0.9
        spiral[k:65 - k, 64 - k] = 0.9; spiral[k + 2:65 - k, k] = 0.9
    _assert_same(spiral, min_area=1)
    checker = (np.indices((64, 64)).sum(0) % 2).astype(np.float32)                 # pure diagonal connectivity
    _assert_same(checker, thresh=0.5, min_area=1)
    dots = np.zeros((256, 256), np.float32)                                        # > max_candidates components
    dots[::3, ::3] = 0.0
    for y in range(0, 256, 6):
        for x in range(0, 256, 6):
            dots[y:y + 4, x:x + 4] = 0.8
    _assert_same(dots, max_candidates=1000, min_area=4)
    _assert_same(dots, max_candidates=37, min_area=4)


def test_rows_wider_than_32_words():
    """Rows of more than 1024 pixels: the row scan of pp_init carries a run from one 32-word segment into the next."""
    rng = np.random.default_rng(5)
    wide = np.zeros((40, 2200), np.float32)
    wide[3, :] = 0.9                       # one run across the whole row (two segment boundaries)
    wide[6, 1000:1060] = 0.9               # crosses x = 1024 inside partly filled words
    wide[9, 992:1088] = 0.9                # whole words on both sides of the boundary
    wide[12, 1024:2048] = 0.9              # a run that is exactly one segment
    wide[15, 0:1024] = 0.9; wide[15, 1025:2200] = 0.9   # ends at a boundary, next run starts one pixel later
    wide[18:22, 2040:2060] = 0.9           # block across x = 2048
    wide[24:, :] = (rng.random((16, 2200)) > 0.35) * 0.9   # dense noise, long ragged runs
    _assert_same(wide, min_area=1, box_thresh=0.5)
    _assert_same(np.full((33, 1100), 0.9, np.float32))      # one component, 35 words per row, 2 strips


def test_full_size_properties():
    """1024x1024, batch 4 (BASELINE config-2 map size): oracle parity + size-independent properties."""
    from oracle import postproc_oracle as ppo
    prob = np.stack([ppo.render_prob_map(1024, 1024, seed=100 + i, n_boxes=180) for i in range(4)])
    labels, ncomp, recs, pp = _run(prob, thresh=0.3, box_thresh=0.5, unclip_ratio=1.6, min_area=10)
    fg = prob > np.float32(0.3)
    for i in range(4):
        assert np.array_equal(labels[i] > 0, fg[i])                                # foreground preserved
        assert labels[i].max() == ncomp[i]
        # idempotence: labelling the label map's foreground again gives the same labels
    relab, _, _, _ = _run((labels > 0).astype(np.float32), thresh=0.5)
    assert np.array_equal(relab, labels)
    # npix of kept boxes sums to no more than the foreground; qsum/npix within [box_thresh, 1]
    for i in range(4):
        assert recs[i]["npix"].sum() <= fg[i].sum()
        s = recs[i]["qsum"].astype(np.float64) / (recs[i]["npix"] * 16777216.0)
        assert (s >= 0.5).all() and (s <= 1.0).all()
    _assert_same(prob[:2], thresh=0.3, box_thresh=0.5, unclip_ratio=1.6, min_area=10)


def test_reference_call_signature_and_cv2_agreement():
    """DBPostProcessor(pred (1,H,W) numpy) -> (list[(K,2) int64 writable], list[float]); agreement with the cv2
    half of the reference path (golden from tests/golden/make_golden.py) is reported as a match rate."""
    from ocr_vi_invoice_b200.postprocess import DBPostProcessor
    from oracle import postproc_oracle as ppo
    p = np.load(GOLD / "postproc_cv2.npz")
    pp = DBPostProcessor(thresh=0.3, box_thresh=0.5, unclip_ratio=1.6)
    pp.min_area = 10
    for i in range(3):
        prob = p[f"{i}.prob"]
        boxes, scores = pp(prob[None])
        assert len(boxes) == len(scores) > 0
        b0 = boxes[0]
        assert b0.shape == (4, 2) and b0.dtype == np.int64 and b0.flags.writeable
        b0[:, 0] = b0[:, 0] / 0.5                                                   # pipeline2.py:325-327
        ours = [(b[:, 0].min(), b[:, 1].min(), b[:, 0].max(), b[:, 1].max()) for b in boxes[1:]]
        theirs = [tuple(int(v) for v in r) for r in p[f"{i}.cv2_rects"]]
        assert ppo.match_rate(ours, theirs, 0.8) >= 0.85
        # torch CUDA input, batched API
        res = pp.batch(torch.from_numpy(prob).cuda()[None, None])
        assert len(res) == 1 and len(res[0][0]) == len(boxes)


def test_gpu_boxes_equal_reference_code_on_rectangular_components():
    """GPU box rectangles == cv2.boundingRect of the boxes returned by the reference's own DBPostProcessor.__call__
    (golden: tests/golden/postproc_reference.npz) for maps whose components are axis-aligned rectangles."""
    g = np.load(GOLD / "postproc_reference.npz")
    checked = 0
    for i in range(int(g["n_maps"])):
        th, bt, ratio, min_area, clean = g[f"params_{i}"]
        if not clean:
            continue
        ref = sorted((int(x), int(y), int(x + w - 1), int(y + h - 1)) for x, y, w, h in g[f"rects_{i}"])
        _, _, recs, pp = _run(g[f"prob_{i}"], thresh=float(th), box_thresh=float(bt), unclip_ratio=float(ratio),
                              min_area=int(min_area))
        boxes, scores = pp.records_to_boxes(recs[0])
        import cv2
        ours = sorted((x, y, x + w - 1, y + h - 1) for x, y, w, h in (cv2.boundingRect(b.astype(np.int32)) for b in boxes))
        assert ours == ref
        assert np.allclose(sorted(scores), sorted(g[f"scores_{i}"]), rtol=0, atol=2.0 ** -23)
        checked += 1
    assert checked >= 2
