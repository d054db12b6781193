"""Drop-in DBNetPP (reference: model/det/dbnet.py:6-17).

Same constructor, same `forward(x) -> dict` with keys binary / thresh / thresh_binary / bin_logits /
thresh_logits (model/det/head.py:42-48), same 754 state_dict keys (strictly loadable from a reference
checkpoint, src/pipeline/pipeline2.py:43-54).  The forward runs the planned sm_100a kernels of engine.py in
bf16 with fp32 accumulation and returns fp32 maps; it refuses CPU tensors and training mode instead of falling
back to eager torch.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from .backbone import ResNet
from .engine import MAP_NAMES, Plan, PreparedWeights
from .head import DBHead
from .neck import FPN_ASF


class DBNetPP(nn.Module):
    def __init__(self, backbone="resnet50", pretrained=True, in_channels=3, inner_channels=256, k=50, dcn=True):
        super().__init__()
        self.backbone = ResNet(name=backbone, pretrained=pretrained, in_channels=in_channels, dcn=dcn)
        self.neck = FPN_ASF(self.backbone.out_channels, inner_channels=inner_channels)
        self.head = DBHead(inner_channels, k=k)
        self._prepared: Optional[PreparedWeights] = None
        self._prepared_sig = None
        self._plans: Dict[Tuple[int, int, int, bool], Plan] = {}
        self.use_cuda_graph = False   # replay a captured graph (outputs then alias plan-owned buffers)
        self.output_maps = MAP_NAMES  # restrict to e.g. ("binary",) to skip writing unused maps

    # ---- weight preparation is invalidated whenever a parameter / buffer is replaced or written in place
    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _get_plan(self, x: torch.Tensor, keep_taps: bool = False) -> Plan:
        u8 = x.dtype == torch.uint8
        shape = (x.shape[0], x.shape[1], x.shape[2]) if u8 else (x.shape[0], x.shape[2], x.shape[3])
        sig = (x.device, self._signature())
        if self._prepared is None or self._prepared_sig != sig:
            self._prepared = PreparedWeights(self, x.device)
            self._prepared_sig = sig
            self._plans.clear()
        key = shape + (keep_taps, u8)
        plan = self._plans.get(key)
        if plan is None:
            plan = Plan(self._prepared, *shape, keep_taps=keep_taps, input_u8=u8)
            self._plans[key] = plan
        return plan

    def _check_input(self, x):
        ok = isinstance(x, torch.Tensor) and x.dim() == 4
        if ok and x.dtype == torch.uint8:
            ok = x.shape[3] == 3  # extension: decoded images, uint8 HWC RGB, normalised on the GPU
        elif ok:
            ok = x.shape[1] == 3 and x.is_floating_point()
        if not ok:
            raise ValueError("DBNetPP.forward expects a float (b, 3, H, W) tensor (or, as an extension, uint8 (b, H, W, 3))")
        if not x.is_cuda:
            raise RuntimeError("B200 DBNet++ path runs on CUDA only: move the model and the input to the GPU "
                               "(there is no CPU fallback)")
        if self.training:
            raise RuntimeError("B200 DBNet++ path is inference-only: call model.eval() first "
                               "(training, model/det/loss.py and src/det/train.py, is out of scope)")
        p = next(self.parameters())
        if p.device != x.device:
            raise RuntimeError(f"model parameters are on {p.device}, input on {x.device}")

    @torch.no_grad()
    def forward(self, x):
        return self.run(x, want=self.output_maps, use_graph=self.use_cuda_graph)

    @torch.no_grad()
    def run(self, x, want=MAP_NAMES, use_graph: bool = False):
        """forward() with the two execution options passed per call (what pipeline.Detector uses, so that wrapping a
        model never changes what a later plain `model(x)` returns): `want` restricts the maps written, `use_graph`
        replays the captured launch list (the outputs then alias plan-owned buffers the next call overwrites)."""
        self._check_input(x)
        with torch.cuda.device(x.device):
            plan = self._get_plan(x)
            xin = x.contiguous() if x.dtype == torch.uint8 else x.float().contiguous()
            return plan.run(xin, want=tuple(want), use_graph=use_graph)

    @torch.no_grad()
    def forward_with_taps(self, x):
        """forward + per-stage activations (NCHW fp32 copies) for the parity tests."""
        self._check_input(x)
        with torch.cuda.device(x.device):
            plan = self._get_plan(x, keep_taps=True)
            out = plan.run(x.float().contiguous(), want=MAP_NAMES)
            return out, {n: plan.tap(n) for n in plan.taps}

    @torch.no_grad()
    def time_launches(self, x, iters: int = 3):
        """Per-kernel-launch device times of the plan for `x` (see Plan.time_launches)."""
        self._check_input(x)
        with torch.cuda.device(x.device):
            plan = self._get_plan(x)
            plan.x_in.copy_(x.float())
            return plan.time_launches(iters)

    def plan_info(self, batch, height, width, device="cuda"):
        """(#launches, #tcgen05 GEMM launches, conv FLOPs) of the plan for an input shape."""
        x = torch.empty(batch, 3, height, width, device=device)
        plan = self._get_plan(x)
        return {"launches": plan.n_launches, "igemm_launches": plan.n_igemm, "flops": plan.flops}
