"""Detection driver around the drop-in model: double-buffered host->device staging, batched GPU post-processing,
and the data-parallel sharding / host gather used on a multi-GPU box.

The reference's production caller (src/pipeline/pipeline2.py:279-328) is a serial batch-1 loop:
H2D -> forward -> D2H of the whole map -> CPU post-processing.  Images are independent, so on N GPUs each rank
(one process per GPU) takes a shard of the image indices, runs the same loop on its own device with no collective,
and rank 0 gathers the variable-length (image_id, boxes, scores) lists on the host.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Sequence, Tuple

import numpy as np
import torch

from .model.det.dbnet import DBNetPP
from .postprocess import DBPostProcessor
from .preprocess import det_size as _det_scales
from .preprocess import resize_image_for_det
from .recprep import RecognitionCropper

Result = Tuple[List[np.ndarray], List[float]]


def shard_indices(n_items: int, rank: int, world: int) -> range:
    """Contiguous shard of `range(n_items)` for `rank`; the first n_items % world ranks get one extra item."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} not in [0, {world})")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_results(local: Sequence[Tuple[int, Result]], group=None, dst: int = 0):
    """Host gather of per-rank [(image_id, (boxes, scores))] lists; returns the merged, id-sorted list on `dst`
    (None elsewhere).  Works on a gloo (CPU) or nccl group; single-process runs return the input sorted."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return sorted(local, key=lambda r: r[0])
    world = dist.get_world_size(group)
    bucket = [None] * world
    dist.all_gather_object(bucket, list(local), group=group)
    if dist.get_rank(group) != dst:
        return None
    merged = [item for part in bucket for item in part]
    ids = [i for i, _ in merged]
    if len(set(ids)) != len(ids):
        raise RuntimeError("gather_results: an image id was produced by more than one rank")
    return sorted(merged, key=lambda r: r[0])


class Detector:
    """model + post-processor with pinned staging and H2D copies overlapped with compute."""

    def __init__(self, model: DBNetPP, post: DBPostProcessor, device=None, use_cuda_graph: bool = True):
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("Detector needs a CUDA device (no CPU fallback)")
        self.model, self.post = model, post
        # per-call options (DBNetPP.run): the caller's module keeps its own use_cuda_graph / output_maps, so a later
        # plain model(x) still returns the five-key dict of the reference forward
        self.use_cuda_graph = bool(use_cuda_graph)
        self.want = ("binary",)  # the box list only needs the probability map (pipeline2.py:318)
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._dev_buf = [None, None]
        self._consumed = [None, None]  # per slot: event after which the main stream no longer reads the slot

    def _stage(self, host: torch.Tensor, slot: int) -> Tuple[torch.Tensor, torch.cuda.Event]:
        """Async H2D of one batch into device slot `slot` on the copy stream."""
        if not host.is_pinned():
            host = host.pin_memory()
        buf = self._dev_buf[slot]
        if buf is None or buf.shape != host.shape or buf.dtype != host.dtype:
            buf = torch.empty(host.shape, dtype=host.dtype, device=self.device)
            self._dev_buf[slot] = buf
        ready = torch.cuda.Event()
        with torch.cuda.stream(self._copy_stream):
            if self._consumed[slot] is not None:
                self._copy_stream.wait_event(self._consumed[slot])
            buf.copy_(host, non_blocking=True)
            ready.record(self._copy_stream)
        return buf, ready

    def detect_and_crop(self, page: torch.Tensor, image_size: int = 640, cropper: RecognitionCropper = None):
        """Steps 2 and 3a of the reference pipeline for one page (pipeline2.py:307-343), without leaving the device:
        resize_image_for_det -> normalise + forward -> post-processing -> box rescale -> crop_image ->
        preprocess_for_recognition.  `page`: the ORIGINAL uint8 HWC CUDA image.  Everything is queued on the current
        stream and nothing synchronises; returns (crop buffers {"out", "rects", "total"}, detector buffers) -- the
        recogniser input is crops["out"][:total] (RecognitionCropper.fetch gives it together with the rectangles)."""
        if cropper is None:
            cropper = self._cropper = getattr(self, "_cropper", None) or RecognitionCropper(device=self.device)
        with torch.cuda.device(self.device):
            resized, scales = resize_image_for_det(page, image_size)
            out = self.model.run(resized[None], want=self.want, use_graph=self.use_cuda_graph)
            det = self.post.run_device(out["binary"][:, 0], want_labels=False)
            crops = cropper.from_detector(det, [page], [scales], int(self.post.max_candidates))
        return crops, det

    def detect(self, host_batch: torch.Tensor) -> List[Result]:
        """One batch: H2D, forward, GPU post-processing, D2H of the box records."""
        return next(self.detect_stream([host_batch]))

    def detect_stream(self, batches: Iterable[torch.Tensor]) -> Iterator[List[Result]]:
        """Yields per-batch results.  The H2D copy of batch i+1 overlaps the kernels of batch i, and the host only
        waits for the box records of batch i after the forward of batch i+1 is queued, so the GPU never idles on the
        host (stream order: fwd(i), post(i), D2H(i) | fwd(i+1), post(i+1), ...)."""
        it = iter(batches)
        try:
            cur = self._stage(next(it), 0)
        except StopIteration:
            return
        slot = 0
        pending = None  # post-processing buffers of the previous batch, D2H queued
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            while cur is not None:
                x, ready = cur
                main.wait_event(ready)
                out = self.model.run(x, want=self.want, use_graph=self.use_cuda_graph)  # copies x into the plan's input buffer, then replays the captured graph
                ev = torch.cuda.Event()
                ev.record(main)
                self._consumed[slot] = ev
                try:
                    cur = self._stage(next(it), slot ^ 1)  # overlaps the forward just queued
                    slot ^= 1
                except StopIteration:
                    cur = None
                if pending is not None:  # the previous batch's records: its kernels ran before this forward
                    recs, _ = self.post.fetch_wait(pending)
                    yield [self.post.records_to_boxes(r) for r in recs]
                pending = self.post.fetch_async(self.post.run_device(out["binary"][:, 0], want_labels=False))
            if pending is not None:
                recs, _ = self.post.fetch_wait(pending)
                yield [self.post.records_to_boxes(r) for r in recs]


class OCRPipeline:
    """Steps 2-3 of the reference's page loop (src/pipeline/pipeline2.py:307-352) on one B200: detection, GPU
    post-processing, box rescale, crop, recognition pre-processing, SVTRv2 and the greedy CTC decode.  Between the page
    upload and the decoded token ids nothing returns to the host except the crop count (to size the recogniser batch)
    and, for the caller's box list, the 48-byte box records.

        ocr = OCRPipeline(det_model, post_processor, rec_model, det_size=640)
        boxes, texts = ocr(page)      # page: ORIGINAL uint8 HWC CUDA image; boxes: int32 (4, 2) arrays in page coordinates
    """

    def __init__(self, det_model: DBNetPP, post: DBPostProcessor, rec_model, device=None, det_size: int = 640,
                 rec_size: Tuple[int, int] = (32, 256), rec_bucket: int = 32, capacity: int = 1024):
        self.detector = Detector(det_model, post, device, use_cuda_graph=False)
        self.rec_model = rec_model  # run with use_graph=True per call: one captured graph per batch bucket, the strings
                                    # are read before the next call (the caller's module attributes are left alone)
        self.det_size = det_size
        self.rec_bucket = int(rec_bucket)
        self.cropper = RecognitionCropper(rec_size, capacity=capacity, device=self.detector.device)

    def __call__(self, page: torch.Tensor):
        det = self.detector
        crops, dbuf = det.detect_and_crop(page, self.det_size, self.cropper)
        det.post.fetch_async(dbuf)
        out, rects = self.cropper.fetch(crops, on_degenerate="zeros")   # the one synchronisation: how many crops
        recs, _ = det.post.fetch_wait(dbuf)
        n = len(rects)
        # pipeline2.py:323-328: boxes back in page coordinates (truncating division), int32
        (_, _), (scale_h, scale_w) = _det_scales(page.shape[0], page.shape[1], self.det_size)
        boxes = []
        for box in det.post.records_to_boxes(recs[0])[0]:
            box[:, 0] = box[:, 0] / scale_w
            box[:, 1] = box[:, 1] / scale_h
            boxes.append(box.astype(np.int32))
        if n == 0:
            return boxes, []
        # the recogniser's plan is keyed by batch size: round the batch up to a bucket (slots past n hold stale crops
        # of the capacity buffer; their strings are dropped)
        nb = min(-(-n // self.rec_bucket) * self.rec_bucket, self.cropper.capacity)
        texts = self.rec_model.decode_probs(self.rec_model.run(crops["out"][:nb], use_graph=True))[:n]
        return boxes, texts
