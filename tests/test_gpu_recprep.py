"""GPU: detection -> recognition hand-over through the C-ABI vs the oracle and the reference goldens.
Bit-exact: rectangles, statuses and every float of the (N, 3, 32, 256) tensors."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).parent / "golden"


def _dev():
    return torch.device("cuda", 0)


def _box_records(boxes, max_candidates):
    """Device image of what b200_postproc writes: only x0, y0, x1, y1 of the 48-byte records matter here."""
    from ocr_vi_invoice_b200.postprocess import _BOX_DTYPE
    rec = np.zeros(max_candidates, dtype=_BOX_DTYPE)
    for i, b in enumerate(boxes):
        rec["x0"][i], rec["y0"][i] = b[:, 0].min(), b[:, 1].min()
        rec["x1"][i], rec["y1"][i] = b[:, 0].max(), b[:, 1].max()
    return rec


def _det_buf(box_lists, max_candidates, dev):
    recs = np.stack([_box_records(b, max_candidates) for b in box_lists])
    return {
        "boxes": torch.from_numpy(recs.view(np.uint8).reshape(-1)).to(dev),
        "counts": torch.tensor([len(b) for b in box_lists], dtype=torch.int32, device=dev),
    }


def test_golden_pages_one_launch():
    """The three reference pages (different sizes and scales) in ONE launch, from device box records."""
    from ocr_vi_invoice_b200.recprep import RecognitionCropper
    g = np.load(GOLD / "recprep_reference.npz")
    dev = _dev()
    names = "abc"
    pages = [torch.from_numpy(g[f"{n}_page"]).to(dev) for n in names]
    scales = [tuple(float(v) for v in g[f"{n}_scales"]) for n in names]
    boxes = [list(g[f"{n}_boxes"]) for n in names]
    cr = RecognitionCropper((32, 256), capacity=256)
    buf = cr.from_detector(_det_buf(boxes, 64, dev), pages, scales, 64)
    out, rects = cr.fetch(buf, on_degenerate="zeros")
    out = out.cpu().numpy()
    table = g["norm_table"]
    at = 0
    for i, n in enumerate(names):
        k = len(boxes[i])
        r = rects[at:at + k]
        assert np.all(r["image"] == i)
        assert np.array_equal(np.stack([r["x"], r["y"], r["w"], r["h"]], axis=1), g[f"{n}_rects"])
        assert np.array_equal(r["status"], g[f"{n}_status"])
        idx = g[f"{n}_tensor_index"]
        want = np.stack([table[c][idx[:, c]] for c in range(3)], axis=1)
        want[g[f"{n}_status"] != 0] = 0.0
        assert np.array_equal(out[at:at + k], want), f"page {n}: tensors differ from the reference's"
        at += k
    assert at == len(rects)
    with pytest.raises(ValueError, match="zero width"):
        cr.fetch(buf)  # default: raise where the reference's cv2.resize raises


def test_random_boxes_against_oracle():
    """Many random boxes (all aspect ratios, borders, outside) on pitched pages, 5 pages per launch, other target size."""
    from oracle import recprep_oracle as ro
    from ocr_vi_invoice_b200.recprep import RecognitionCropper
    dev = _dev()
    rng = np.random.default_rng(5)
    pages_np, pages, scales, boxes = [], [], [], []
    for i in range(5):
        h, w = int(rng.integers(40, 700)), int(rng.integers(40, 900))
        pg = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        pages_np.append(pg)
        wide = torch.zeros(h, w + 7, 3, dtype=torch.uint8, device=dev)  # row pitch != 3 * w
        wide[:, :w] = torch.from_numpy(pg).to(dev)
        pages.append(wide[:, :w])
        sh, sw = float(rng.uniform(0.4, 2.5)), float(rng.uniform(0.4, 2.5))
        scales.append((sh, sw))
        bl = []
        for _ in range(int(rng.integers(1, 120))):
            x0, y0 = int(rng.integers(-30, int(w * sw) + 20)), int(rng.integers(-30, int(h * sh) + 20))
            bw, bh = int(rng.integers(0, 500)), int(rng.integers(0, 90))
            bl.append(np.array([[x0, y0], [x0 + bw, y0], [x0 + bw, y0 + bh], [x0, y0 + bh]], dtype=np.int64))
        boxes.append(bl)
    for size in ((32, 256), (48, 320), (20, 100)):
        cr = RecognitionCropper(size, capacity=640)
        out, rects = cr.fetch(cr.from_detector(_det_buf(boxes, 128, dev), pages, scales, 128), on_degenerate="zeros")
        out = out.cpu().numpy()
        at = 0
        for i in range(5):
            want, wr = ro.crops_for_recognition(pages_np[i], boxes[i], scales[i][0], scales[i][1], size)
            k = len(boxes[i])
            r = rects[at:at + k]
            assert np.array_equal(np.stack([r["x"], r["y"], r["w"], r["h"], r["status"]], axis=1), wr)
            assert np.array_equal(out[at:at + k], want)
            at += k
        assert at == len(rects)


def test_capacity_and_empty_batch():
    from ocr_vi_invoice_b200.recprep import RecognitionCropper
    dev = _dev()
    page = torch.full((64, 64, 3), 200, dtype=torch.uint8, device=dev)
    box = np.array([[4, 4], [40, 4], [40, 20], [4, 20]], dtype=np.int64)
    cr = RecognitionCropper((32, 256), capacity=4)
    out, rects = cr.fetch(cr.from_detector(_det_buf([[]], 8, dev), [page], [(1.0, 1.0)], 8))
    assert out.shape[0] == 0 and len(rects) == 0
    with pytest.raises(RuntimeError, match="capacity"):
        cr.fetch(cr.from_detector(_det_buf([[box] * 6], 8, dev), [page], [(1.0, 1.0)], 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cr.from_detector(_det_buf([[box]], 8, dev), [page.cpu()], [(1.0, 1.0)], 8)


def test_reference_named_functions():
    """crop_image / preprocess_for_recognition / recognition_batch with the reference's names, on CUDA tensors."""
    from oracle import recprep_oracle as ro
    from ocr_vi_invoice_b200 import recprep
    g = np.load(GOLD / "recprep_reference.npz")
    dev = _dev()
    page_np = g["a_page"]
    page = torch.from_numpy(page_np).to(dev)
    sh, sw = (float(v) for v in g["a_scales"])
    crops, want = [], []
    for box, st in zip(g["a_boxes"], g["a_status"]):
        rb = ro.rescale_box(box, sh, sw)
        crop = recprep.crop_image(page, rb)
        ref_crop = ro.crop_image(page_np, rb)
        assert tuple(crop.shape[:2]) == tuple(ref_crop.shape[:2]) or (crop.numel() == 0 and ref_crop.size == 0)
        if st == ro.STATUS_DEGENERATE:
            with pytest.raises(recprep.DegenerateCropError):
                recprep.preprocess_for_recognition(crop)
            continue
        if st == ro.STATUS_OK:
            assert np.array_equal(crop.cpu().numpy(), ref_crop)
            t = recprep.preprocess_for_recognition(crop)
            assert np.array_equal(t.cpu().numpy(), ro.preprocess_for_recognition(ref_crop)[0])
        crops.append(crop)
        want.append(ro.preprocess_for_recognition(ref_crop)[0])
    batch = recprep.recognition_batch(crops)
    assert np.array_equal(batch.cpu().numpy(), np.stack(want))


def test_detector_to_recogniser_without_host_round_trip():
    """DBPostProcessor.run_device -> RecognitionCropper.from_detector on the same stream: the crops equal the oracle's
    run on the host copies of the very same box list."""
    from oracle import recprep_oracle as ro
    from ocr_vi_invoice_b200.postprocess import DBPostProcessor
    from ocr_vi_invoice_b200.preprocess import det_size
    from ocr_vi_invoice_b200.recprep import RecognitionCropper
    dev = _dev()
    rng = np.random.default_rng(9)
    pages_np, probs, scales = [], [], []
    H = W = 256
    for i in range(3):
        h, w = int(rng.integers(300, 500)), int(rng.integers(300, 500))
        pages_np.append(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8))
        scales.append((H / h, W / w))
        prob = np.full((H, W), 0.05, dtype=np.float32)
        for _ in range(12):
            bh, bw = int(rng.integers(6, 20)), int(rng.integers(10, 90))
            y, x = int(rng.integers(0, H - bh)), int(rng.integers(0, W - bw))
            prob[y:y + bh, x:x + bw] = 0.9
        probs.append(prob)
    post = DBPostProcessor(thresh=0.3, box_thresh=0.5, max_candidates=100, device=dev)
    det = post.run_device(torch.from_numpy(np.stack(probs)).to(dev), want_labels=False)
    cr = RecognitionCropper((32, 256), capacity=128)
    buf = cr.from_detector(det, [torch.from_numpy(p).to(dev) for p in pages_np], scales, post.max_candidates)
    out, rects = cr.fetch(buf, on_degenerate="zeros")
    recs, _ = post.fetch(det)
    at = 0
    for i in range(3):
        boxes, _ = post.records_to_boxes(recs[i])
        assert len(boxes) > 0
        want, wr = ro.crops_for_recognition(pages_np[i], boxes, scales[i][0], scales[i][1])
        k = len(boxes)
        assert np.array_equal(rects["status"][at:at + k], wr[:, 4])
        assert np.array_equal(out[at:at + k].cpu().numpy(), want)
        at += k
    assert at == len(rects)
    assert det_size(480, 640, 640)[0] == (480, 640)


def test_page_to_crops_chain_matches_reference_chain():
    """Detector.detect_and_crop (GPU resize -> forward -> post-processing -> crops, nothing returns to the host) vs the
    oracle chain run on the SAME probability map: resize restatement, post-processing oracle, crop oracle."""
    from oracle import postproc_oracle as ppo
    from oracle import preprocess_oracle as pre
    from oracle import recprep_oracle as ro
    from oracle.weights import make_state_dict
    from ocr_vi_invoice_b200.model.det.dbnet import DBNetPP
    from ocr_vi_invoice_b200.pipeline import Detector
    from ocr_vi_invoice_b200.postprocess import DBPostProcessor
    from ocr_vi_invoice_b200.recprep import RecognitionCropper
    from ocr_vi_invoice_b200.synth import synthetic_batch_u8
    dev = _dev()
    model = DBNetPP(pretrained=False)
    model.load_state_dict(make_state_dict(seed=3, offset_px=0.5), strict=True)
    model = model.to(dev).eval()
    post = DBPostProcessor(thresh=0.3, box_thresh=0.5, unclip_ratio=1.6, device=dev)
    det = Detector(model, post, dev, use_cuda_graph=False)
    page_np = np.ascontiguousarray(synthetic_batch_u8(1, 300, 420, seed=4)[0])
    page = torch.from_numpy(page_np).to(dev)
    cr = RecognitionCropper((32, 256), capacity=1024)
    crops, dbuf = det.detect_and_crop(page, image_size=256, cropper=cr)
    out, rects = cr.fetch(crops, on_degenerate="zeros")
    # the reference chain on the host, from the very same probability map
    resized, (sh, sw) = pre.resize_image_for_det(page_np, 256)
    prob = model(torch.from_numpy(resized).to(dev)[None])["binary"][0, 0].cpu().numpy()
    _, _, recs = ppo.postprocess(prob, 0.3, 0.5, 1000, 1.6, 10)
    boxes = [np.array([[r["x0"], r["y0"]], [r["x1"], r["y0"]], [r["x1"], r["y1"]], [r["x0"], r["y1"]]], dtype=np.int64)
             for r in recs]
    want, wr = ro.crops_for_recognition(page_np, boxes, sh, sw)
    assert len(rects) == len(boxes)
    assert np.array_equal(np.stack([rects["x"], rects["y"], rects["w"], rects["h"], rects["status"]], axis=1).reshape(-1, 5),
                          wr.reshape(-1, 5))
    assert np.array_equal(out.cpu().numpy(), want)
