"""Drop-in DBPostProcessor (reference: src/det/test.py:46-106) running on the GPU.

Same constructor arguments and attributes (`thresh`, `box_thresh`, `max_candidates`, `unclip_ratio`, `min_size`,
`min_area`), same call: `boxes, scores = post(pred)` with `pred` a (1, H, W) float32 map returning
`list[np.ndarray (K, 2) int64]`, `list[float]` -- writable arrays, because the caller rescales them in place
(src/pipeline/pipeline2.py:323-328).  Differences from the reference, by design of the north star:
contours/Clipper are replaced by connected components + a closed-form rectangle un-clip, so K == 4 and boxes
come in raster order of their first pixel.  `pred` may also be a CUDA tensor, (1,H,W) / (B,H,W) / (B,1,H,W),
which avoids the full-map D2H copy the reference does at pipeline2.py:320; only the box records come back.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from . import _lib, ops

Q = 16777216.0
_BOX_DTYPE = np.dtype([("x0", "<i4"), ("y0", "<i4"), ("x1", "<i4"), ("y1", "<i4"), ("bx0", "<i4"), ("by0", "<i4"),
                       ("bx1", "<i4"), ("by1", "<i4"), ("label", "<i4"), ("npix", "<i4"), ("qsum", "<u8")])
assert _BOX_DTYPE.itemsize == 48


class DBPostProcessor:
    def __init__(self, thresh=0.3, box_thresh=0.6, max_candidates=1000, unclip_ratio=1.5, device=None):
        self.thresh = thresh
        self.box_thresh = box_thresh
        self.max_candidates = max_candidates
        self.unclip_ratio = unclip_ratio
        self.min_size = 3    # unused by the reference too (test.py:52)
        self.min_area = 10   # callers overwrite it after construction (pipeline2.py:259)
        self.device = device
        self._bufs = {}

    # ------------------------------------------------------------------ device side
    def _buffers(self, b, h, w, dev):
        key = (b, h, w, int(self.max_candidates), str(dev))
        buf = self._bufs.get(key)
        if buf is None:
            mc = int(self.max_candidates)
            ws = torch.empty(ops.postproc_workspace_bytes(b, h, w, mc), dtype=torch.uint8, device=dev)
            buf = {
                "ws": ws,
                "labels": torch.empty(b, h, w, dtype=torch.int32, device=dev),
                "boxes": torch.empty(b * mc * 48, dtype=torch.uint8, device=dev),
                "counts": torch.empty(b, dtype=torch.int32, device=dev),
                "ncomp": torch.empty(b, dtype=torch.int32, device=dev),
                "host_boxes": torch.empty(b * mc * 48, dtype=torch.uint8).pin_memory(),
                "host_counts": torch.empty(2 * b, dtype=torch.int32).pin_memory(),
            }
            self._bufs = {key: buf}  # keep one shape cached
        return buf

    def run_device(self, prob: torch.Tensor, want_labels: bool = True):
        """prob: CUDA float32 (B, H, W).  Launches the kernels on the current stream; returns the buffer dict
        (device `labels`, `boxes`, `counts`, `ncomp`).  No synchronisation."""
        if not prob.is_cuda:
            raise RuntimeError("DBPostProcessor.run_device needs a CUDA tensor")
        prob = prob.contiguous().float()
        b, h, w = prob.shape
        buf = self._buffers(b, h, w, prob.device)
        with torch.cuda.device(prob.device):
            ops.postproc(prob, buf["labels"] if want_labels else None, buf["boxes"], buf["counts"], buf["ncomp"],
                         buf["ws"], b, h, w, self.thresh, self.box_thresh, self.unclip_ratio, self.min_area,
                         self.max_candidates)
        buf["prob_ref"] = prob  # keep alive until the stream has consumed it
        return buf

    def fetch_async(self, buf):
        """Queues the D2H of the box records (48 B per slot) behind the kernels of `run_device` and records an event;
        the caller may queue more GPU work before `fetch_wait`."""
        b = buf["counts"].numel()
        hc = buf["host_counts"]
        hc[:b].copy_(buf["counts"], non_blocking=True)
        hc[b:].copy_(buf["ncomp"], non_blocking=True)
        buf["host_boxes"].copy_(buf["boxes"], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        buf["fetch_event"] = ev
        return buf

    def fetch_wait(self, buf) -> Tuple[List[np.ndarray], np.ndarray]:
        """Waits for `fetch_async` and returns (records per image, ncomp)."""
        buf["fetch_event"].synchronize()
        b = buf["counts"].numel()
        mc = int(self.max_candidates)
        hc = buf["host_counts"]
        counts = hc[:b].numpy()
        recs_all = buf["host_boxes"].numpy().view(_BOX_DTYPE).reshape(b, mc)
        return [recs_all[i, :counts[i]].copy() for i in range(b)], hc[b:].numpy().copy()

    def fetch(self, buf) -> Tuple[List[np.ndarray], np.ndarray]:
        """D2H of the box records only (48 B per kept box).  Returns (records per image, ncomp)."""
        return self.fetch_wait(self.fetch_async(buf))

    @staticmethod
    def records_to_boxes(recs) -> Tuple[List[np.ndarray], List[float]]:
        boxes = [np.array([[r["x0"], r["y0"]], [r["x1"], r["y0"]], [r["x1"], r["y1"]], [r["x0"], r["y1"]]],
                          dtype=np.int64) for r in recs]
        scores = [float(np.float64(r["qsum"]) / (np.float64(r["npix"]) * Q)) for r in recs]
        return boxes, scores

    # ------------------------------------------------------------------ reference-compatible entry points
    def _to_device_maps(self, pred):
        if isinstance(pred, np.ndarray):
            dev = self.device or "cuda"
            pred = torch.from_numpy(np.ascontiguousarray(pred, dtype=np.float32)).to(dev, non_blocking=True)
        if pred.dim() == 4:
            pred = pred[:, 0]
        elif pred.dim() == 2:
            pred = pred[None]
        return pred

    def __call__(self, pred, is_output_polygon=False):
        """pred: (C, H, W) with C == 1 (test.py:56) -- numpy float32 or CUDA tensor.  Returns (boxes, scores)."""
        maps = self._to_device_maps(pred)
        recs, _ = self.fetch(self.run_device(maps[:1], want_labels=False))
        return self.records_to_boxes(recs[0])

    def batch(self, pred) -> List[Tuple[List[np.ndarray], List[float]]]:
        """All images of a (B, 1, H, W) / (B, H, W) CUDA map in one launch sequence + one small D2H."""
        maps = self._to_device_maps(pred)
        recs, _ = self.fetch(self.run_device(maps, want_labels=False))
        return [self.records_to_boxes(r) for r in recs]
