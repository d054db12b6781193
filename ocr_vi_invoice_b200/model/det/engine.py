"""Execution plan of the DBNet++ forward on one B200: weight preparation + a static list of kernel launches.

The reference runs ~250 eager torch ops per forward.  Here a plan is built once per
(batch, H, W): every convolution becomes one tcgen05 implicit-GEMM launch with BN / bias / ReLU / residual /
nearest-upsample-add fused in its epilogue, activations live in NHWC bf16 buffers owned by the plan, and the
launch list can be captured into a CUDA graph.  Layer order follows model/det/backbone.py:55-60,
torchvision Bottleneck.forward (resnet.py:143-163), model/det/neck.py:26-46,57-79 and model/det/head.py:32-48.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ... import ops

LAYER_CFG = ((1, 3, 64), (2, 4, 128), (3, 6, 256), (4, 3, 512))
MAP_NAMES = ("binary", "thresh", "thresh_binary", "bin_logits", "thresh_logits")


def _bn_fold(bn, conv_bias=None):
    return ops.fold_bn(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps, conv_bias)


class PreparedWeights:
    """bf16 K-major weight matrices + fp32 per-column scale/shift for every layer (built once per state)."""

    def __init__(self, model, device):
        bb = model.backbone.model
        self.device = device
        self.dcn = model.backbone.dcn
        to = lambda t: t.detach().to(device=device, dtype=torch.float32)
        P: Dict[str, tuple] = {}

        def conv_bn(name, conv_w, bn, n_pad=None, conv_bias=None):
            w = ops.pack_conv_weight(to(conv_w), n_pad)
            s, t = _bn_fold(bn, None if conv_bias is None else to(conv_bias))
            P[name] = (w, ops.pad_vec(to(s), w.shape[0], 1.0), ops.pad_vec(to(t), w.shape[0]))

        # stem (torchvision resnet.py:197-199)
        w = ops.pack_stem_weight(to(bb.conv1.weight), 64)
        s, t = _bn_fold(bb.bn1)
        P["stem"] = (w, ops.pad_vec(to(s), 64, 1.0), ops.pad_vec(to(t), 64))
        for li, nblocks, _width in LAYER_CFG:
            layer = getattr(bb, f"layer{li}")
            for bi in range(nblocks):
                blk = layer[bi]
                p = f"layer{li}.{bi}"
                conv_bn(p + ".conv1", blk.conv1.weight, blk.bn1)
                if self.dcn and li >= 2:
                    om = blk.conv2.offset_mask_conv
                    w = ops.pack_conv_weight(to(om.weight), 32)
                    P[p + ".offset"] = (w, torch.ones(32, device=device), ops.pad_vec(to(om.bias), 32))
                if self.dcn and li >= 2 and ops.dcn_tmem_supported(blk.conv2.weight.shape[0]):
                    # c3 / c4: tensor-memory DCN kernel, weights in its K order
                    w = ops.pack_dcn_weight(to(blk.conv2.weight))
                    s, t = _bn_fold(blk.bn2)
                    P[p + ".conv2"] = (w, ops.pad_vec(to(s), w.shape[0], 1.0), ops.pad_vec(to(t), w.shape[0]))
                else:
                    conv_bn(p + ".conv2", blk.conv2.weight, blk.bn2)
                conv_bn(p + ".conv3", blk.conv3.weight, blk.bn3)
                if blk.downsample is not None:
                    conv_bn(p + ".down", blk.downsample[0].weight, blk.downsample[1])
                    # The projection shortcut inside conv3 (resnet.py:155-161): bn3(conv3(t)) + bn_d(conv_d(x)) is ONE GEMM over
                    # K = [t | x] once both BN scales are folded into the weights (fp32, before the bf16 rounding); the
                    # shifts add up.  The launch reads x as a second K segment (b200_igemm_desc.x2).
                    s3, t3 = _bn_fold(blk.bn3)
                    sd, td = _bn_fold(blk.downsample[1])
                    w3 = ops.pack_conv_weight(to(blk.conv3.weight) * to(s3).view(-1, 1, 1, 1))
                    wd = ops.pack_conv_weight(to(blk.downsample[0].weight) * to(sd).view(-1, 1, 1, 1), w3.shape[0])
                    wf = torch.cat([w3, wd], dim=1).contiguous()
                    P[p + ".conv3d"] = (wf, torch.ones(wf.shape[0], device=device), ops.pad_vec(to(t3) + to(td), wf.shape[0]))
        nk = model.neck
        for i in range(4):
            lc = nk.lateral_convs[i]
            w = ops.pack_conv_weight(to(lc.weight))
            P[f"lat{i}"] = (w, torch.ones(w.shape[0], device=device), ops.pad_vec(to(lc.bias), w.shape[0]))
            conv_bn(f"smooth{i}", nk.fpn_convs[i].conv.weight, nk.fpn_convs[i].bn)
        self.w_att = to(nk.asf.conv_atten.weight).reshape(4, 1024).contiguous()
        self.b_att = to(nk.asf.conv_atten.bias).contiguous()
        for i in range(4):  # z_l = W_l p_l: the level's 256 attention columns as a 1x1 conv with N padded to 32
            wz = ops.pack_conv_weight(self.w_att[:, 256 * i:256 * (i + 1)].reshape(4, 256, 1, 1), 32)
            P[f"asf.z{i}"] = (wz, torch.ones(32, device=device), torch.zeros(32, device=device))
        # head: the first 3x3 conv of both branches is one GEMM (N = 64 + 64, reads the fused map once)
        hb, ht = model.head.bin_conv, model.head.thresh_conv
        w0 = torch.cat([ops.pack_conv_weight(to(hb[0].conv.weight), 64), ops.pack_conv_weight(to(ht[0].conv.weight), 64)])
        sb, tb = _bn_fold(hb[0].bn)
        st, tt = _bn_fold(ht[0].bn)
        P["head.0"] = (w0.contiguous(), torch.cat([to(sb), to(st)]).contiguous(), torch.cat([to(tb), to(tt)]).contiguous())
        for br, seq in (("bin", hb), ("thr", ht)):
            # BN scale folded into the transposed conv's weights (fp32, before the bf16 rounding): the fused head-tail
            # epilogue adds the shift inside the UMMA chain and fetches no per-column constants (B200_EPI_HEAD_DOT)
            s, t = _bn_fold(seq[2], to(seq[1].bias))
            w = ops.pack_convT_weight(to(seq[1].weight) * to(s).view(1, -1, 1, 1), 256)
            P[f"head.{br}.1"] = (w, torch.ones(256, device=device), to(t).repeat(4).contiguous())
            # last ConvTranspose2d(64 -> 1): [ci][dy*2+dx] fp32, consumed by the head-tail kernel
            P[f"head.{br}.4"] = (to(seq[4].weight).reshape(64, 4).contiguous(), float(seq[4].bias.detach().float().item()))
        self.P = P
        self.k = float(model.head.k)


class Plan:
    """Static launch list + activation buffers for one (batch, H, W)."""

    def __init__(self, W: PreparedWeights, batch: int, H: int, Wd: int, keep_taps: bool = False, input_u8: bool = False):
        if H % 32 or Wd % 32:
            raise ValueError(f"DBNet++ input sizes must be multiples of 32 (got {H}x{Wd}); see pipeline2.py:33-40")
        dev = W.device
        self.W, self.batch, self.H, self.Wd = W, batch, H, Wd
        self.steps: List = []
        self.names: List[str] = []
        self.taps: Dict[str, tuple] = {}  # name -> (NHWC tensor, channels)
        self.flops = 0.0
        self.n_igemm = 0
        self.step_flops: Dict[str, float] = {}
        self.step_kind: Dict[str, str] = {}
        self.step_bytes: Dict[str, float] = {}  # algorithmic HBM bytes: every operand read once, every output written once
        self.step_write_bytes: Dict[str, float] = {}  # the written part of step_bytes (HBM writes top out below the copy rate)
        b = batch
        bf = lambda *s: torch.empty(*s, device=dev, dtype=torch.bfloat16)
        P = W.P

        def igemm(name, mode, x, key, y, *, in_h, in_w, in_c, out_c, ksize, stride, flags=0, residual=None, res_ld=0,
                  offsets=None, shuf_c=0, out_ld=None, x_ld=0, chain=None, x2=None, in_c2=0, stride2=1):
            w, s, t = P[key]
            op = ops.Igemm(mode, x, w, s, t, y, batch=b, in_h=in_h, in_w=in_w, in_c=in_c, out_c=out_c, ksize=ksize,
                           stride=stride, flags=flags, residual=residual, res_ld=res_ld, offsets=offsets, shuf_c=shuf_c,
                           out_ld=out_ld, x_ld=x_ld, name=name, chain=chain, x2=x2, in_c2=in_c2, stride2=stride2)
            self.steps.append(op)
            self.names.append(name)
            self.step_flops[name] = op.flops
            self.step_kind[name] = {ops.IGEMM_CONV: "conv", ops.IGEMM_STEM: "stem", ops.IGEMM_DCN: "dcn",
                                    ops.IGEMM_DCN_TMEM: "dcn"}[mode]
            oh, ow = in_h // stride, in_w // stride
            out_px = b * oh * ow * (4 if flags & ops.EPI_SHUFFLE2 else 1)
            out_bytes = out_px * 4 * 4 if flags & ops.EPI_HEAD_DOT else y.element_size() * b * oh * ow * out_c
            in_px = b * in_h * in_w if ksize > 1 or stride == 1 else b * oh * ow  # a strided 1x1 reads only its pixels
            res_bytes = (2 * b * oh * ow * out_c if flags & ops.EPI_RES else 0) + (2 * b * oh * ow * out_c // 4 if flags & ops.EPI_RES_UP2 else 0)
            self.step_bytes[name] = float(in_px * in_c * 2 + out_bytes + res_bytes + w.numel() * 2)
            if x2 is not None:
                self.step_bytes[name] += float(b * oh * ow * in_c2 * 2)  # the second K segment reads only the pixels it samples
            self.step_write_bytes[name] = float(out_bytes)
            if chain is not None:
                self.step_bytes[name] += float(chain[3].numel() * 2 + chain[0].numel() * 2)
                self.step_write_bytes[name] += float(chain[3].numel() * 2)
            self.flops += op.flops
            self.n_igemm += 1
            return op

        # ---- stem: pack -> 7x7 s2 (+BN+ReLU) -> maxpool
        packed = bf(b, H // 2 + 3, Wd // 2 + 3, 16)
        stem = bf(b, H // 2, Wd // 2, 64)
        if input_u8:  # decoded images, uint8 HWC RGB: normalisation fused into the packing kernel
            self.x_in = torch.empty(b, H, Wd, 3, device=dev, dtype=torch.uint8)
            self.steps.append(lambda: ops.stem_pack_u8(self.x_in, packed)); self.names.append("stem_pack_u8")
        else:         # the reference's interface: normalised fp32 NCHW
            self.x_in = torch.empty(b, 3, H, Wd, device=dev, dtype=torch.float32)
            self.steps.append(lambda: ops.stem_pack(self.x_in, packed)); self.names.append("stem_pack")
        igemm("stem", ops.IGEMM_STEM, packed, "stem", stem, in_h=H, in_w=Wd, in_c=3, out_c=64, ksize=7, stride=2,
              flags=ops.EPI_RELU)
        h, w = H // 4, Wd // 4
        x = bf(b, h, w, 64)
        self.steps.append(lambda s=stem, o=x: ops.maxpool3x3s2(s, o, b, H // 2, Wd // 2, 64)); self.names.append("maxpool")
        cin = 64
        feats = []
        import os
        use_chain = not os.environ.get("B200_NO_CHAIN")
        use_fused_down = not os.environ.get("B200_NO_FUSED_DOWN")
        chained_t1 = None  # conv1 output of this block, already produced by the previous block's conv3 launch
        for li, nblocks, width in LAYER_CFG:
            cout = width * 4
            for bi in range(nblocks):
                p = f"layer{li}.{bi}"
                stride = 2 if (bi == 0 and li > 1) else 1
                ho, wo = h // stride, w // stride
                if chained_t1 is not None:
                    t1, chained_t1 = chained_t1, None
                else:
                    t1 = bf(b, h, w, width)
                    igemm(p + ".conv1", ops.IGEMM_CONV, x, p + ".conv1", t1, in_h=h, in_w=w, in_c=cin, out_c=width, ksize=1,
                          stride=1, flags=ops.EPI_RELU)
                t2 = bf(b, ho, wo, width)
                if W.dcn and li >= 2:
                    # offset_mask_conv with the sampling-record epilogue: [pixel][9 taps][16-byte record]
                    om = torch.empty(b, ho, wo, 9, 4, device=dev, dtype=torch.int32)
                    igemm(p + ".offset", ops.IGEMM_CONV, t1, p + ".offset", om, in_h=h, in_w=w, in_c=width, out_c=32,
                          ksize=3, stride=stride, flags=ops.EPI_DCN_COORDS, out_ld=32)
                    igemm(p + ".dcn", ops.IGEMM_DCN_TMEM if ops.dcn_tmem_supported(width) else ops.IGEMM_DCN, t1, p + ".conv2", t2,
                          in_h=h, in_w=w, in_c=width, out_c=width, ksize=3, stride=stride, flags=ops.EPI_RELU, offsets=om)
                else:
                    igemm(p + ".conv2", ops.IGEMM_CONV, t1, p + ".conv2", t2, in_h=h, in_w=w, in_c=width, out_c=width,
                          ksize=3, stride=stride, flags=ops.EPI_RELU)
                fused_down = (bi == 0 and use_fused_down)
                if fused_down:
                    idn = None  # the projection shortcut runs inside conv3 (second K segment)
                elif bi == 0:
                    idn = bf(b, ho, wo, cout)
                    igemm(p + ".down", ops.IGEMM_CONV, x, p + ".down", idn, in_h=h, in_w=w, in_c=cin, out_c=cout,
                          ksize=1, stride=stride)
                else:
                    idn = x
                y = bf(b, ho, wo, cout)
                chain = None
                if use_chain and cout == 256:
                    # layer1: the next block's conv1 (256 -> 64, or layer2.0's 256 -> 128) + bn1 + ReLU rides on this launch's
                    # staged output chunks (B200_EPI_CHAIN): the 256-channel block output is not read back by a conv1 launch
                    nxt = f"layer{li}.{bi + 1}.conv1" if bi + 1 < nblocks else f"layer{li + 1}.0.conv1"
                    w2, s2, t2_ = P[nxt]
                    chained_t1 = bf(b, ho, wo, w2.shape[0])
                    chain = (w2, s2, t2_, chained_t1, True)
                if fused_down:
                    igemm(p + ".conv3", ops.IGEMM_CONV, t2, p + ".conv3d", y, in_h=ho, in_w=wo, in_c=width, out_c=cout,
                          ksize=1, stride=1, flags=ops.EPI_RELU, chain=chain, x2=x, in_c2=cin, stride2=stride)
                else:
                    igemm(p + ".conv3", ops.IGEMM_CONV, t2, p + ".conv3", y, in_h=ho, in_w=wo, in_c=width, out_c=cout,
                          ksize=1, stride=1, flags=ops.EPI_RES | ops.EPI_RELU, residual=idn, res_ld=cout, chain=chain)
                if keep_taps and W.dcn and li >= 2:
                    self.taps[f"backbone.model.{p}.bn2"] = (t2, width)
                x, cin, h, w = y, cout, ho, wo
            feats.append((x, cin, h, w))
            if keep_taps:
                self.taps[f"c{li + 1}"] = (x, cin)
        # ---- neck: lateral 1x1 (+bias) with the top-down nearest-upsample add fused, then 3x3 smoothing
        inner = [None] * 4
        ps = [None] * 4
        for i in (3, 2, 1, 0):
            f, c, fh, fw = feats[i]
            inner[i] = bf(b, fh, fw, 256)
            if i == 3:
                igemm(f"lat{i}", ops.IGEMM_CONV, f, f"lat{i}", inner[i], in_h=fh, in_w=fw, in_c=c, out_c=256, ksize=1,
                      stride=1)
            else:
                igemm(f"lat{i}", ops.IGEMM_CONV, f, f"lat{i}", inner[i], in_h=fh, in_w=fw, in_c=c, out_c=256, ksize=1,
                      stride=1, flags=ops.EPI_RES_UP2, residual=inner[i + 1], res_ld=256)
            ps[i] = bf(b, fh, fw, 256)
            igemm(f"smooth{i}", ops.IGEMM_CONV, inner[i], f"smooth{i}", ps[i], in_h=fh, in_w=fw, in_c=256, out_c=256,
                  ksize=3, stride=1, flags=ops.EPI_RELU)
            if keep_taps:
                self.taps[f"p{i + 2}"] = (ps[i], 256)
        sizes = [(feats[i][2], feats[i][3]) for i in range(4)]
        fused = bf(b, sizes[0][0], sizes[0][1], 256)
        zs = []
        for i in range(4):  # attention-logit contributions at native resolution (fp32, 8 floats per pixel, 4 used)
            z = torch.empty(b, sizes[i][0], sizes[i][1], 8, device=dev, dtype=torch.float32)
            igemm(f"asf.z{i}", ops.IGEMM_CONV, ps[i], f"asf.z{i}", z, in_h=sizes[i][0], in_w=sizes[i][1], in_c=256, out_c=8,
                  ksize=1, stride=1, flags=ops.EPI_OUT_F32, out_ld=8)
            zs.append(z)
        if ops.AsfPlan.supported(sizes):  # bilinear upsampling + attention-weighted sum as one UMMA chain per tile
            self.steps.append(ops.AsfPlan(ps, zs, 8, W.b_att, fused, b, sizes))
        else:                             # tiny maps (a level smaller than its gather window): CUDA-core blend
            self.steps.append(lambda: ops.asf_combine(ps, zs, 8, W.b_att, fused, b, sizes))
        self.names.append("asf")
        if keep_taps:
            self.taps["asf"] = (fused, 256)
        # ---- head
        h4, w4 = sizes[0]
        tails = {}
        t01 = bf(b, h4, w4, 128)  # [bin 64 | thr 64] channels per pixel
        igemm("head.0", ops.IGEMM_CONV, fused, "head.0", t01, in_h=h4, in_w=w4, in_c=256, out_c=128, ksize=3, stride=1,
              flags=ops.EPI_RELU)
        self.maps: Optional[Dict[str, torch.Tensor]] = None
        self.want = MAP_NAMES
        self.graph = None
        if keep_taps:
            # unfused tail (tests read the 64-channel maps): ConvT -> bf16 map -> db_head kernel
            for bi_, br in enumerate(("bin", "thr")):
                u = bf(b, 2 * h4, 2 * w4, 64)
                igemm(f"head.{br}.1", ops.IGEMM_CONV, t01[..., 64 * bi_:], f"head.{br}.1", u, in_h=h4, in_w=w4, in_c=64,
                      out_c=256, ksize=1, stride=1, flags=ops.EPI_RELU | ops.EPI_SHUFFLE2, shuf_c=64, x_ld=128)
                tails[br] = u
                self.taps[f"head.{'bin_conv' if br == 'bin' else 'thresh_conv'}.3"] = (u, 64)
            self.tails = tails

            def head_tail():
                (wb, bb_), (wt, bt) = P["head.bin.4"], P["head.thr.4"]
                ops.db_head_fwd(tails["bin"], tails["thr"], wb, bb_, wt, bt, W.k, self.maps, b, 2 * h4, 2 * w4)

            self.steps.append(head_tail)
            self.names.append("db_head")
        else:
            # fused tail: the first transposed conv's epilogue applies BN + ReLU + bf16 rounding + the last transposed
            # conv (64 -> 1) and writes the fp32 logit map; the 64-channel maps (2 x 537 MB at batch 16) never exist.
            # Bit-identical to the unfused tail (same channel order and fmaf sequence).
            self.logits = {}
            self._tail_w = {}
            for bi_, br in enumerate(("bin", "thr")):
                w2, b2 = P[f"head.{br}.4"]
                aux = torch.cat([w2.reshape(-1).float(), torch.tensor([b2], device=dev, dtype=torch.float32)]).contiguous()
                self._tail_w[br] = aux
                lg = torch.empty(b, 1, 4 * h4, 4 * w4, device=dev, dtype=torch.float32)
                self.logits[br] = lg
                igemm(f"head.{br}.1", ops.IGEMM_CONV, t01[..., 64 * bi_:], f"head.{br}.1", lg, in_h=h4, in_w=w4, in_c=64,
                      out_c=256, ksize=1, stride=1, flags=ops.EPI_RELU | ops.EPI_SHUFFLE2 | ops.EPI_HEAD_DOT, shuf_c=64,
                      x_ld=128, offsets=aux)
            self.steps.append(lambda: ops.db_maps(self.logits["bin"], self.logits["thr"], W.k, self.maps))
            self.names.append("db_maps")
        self.n_launches = len(self.steps)

    def _alloc_maps(self):
        dev = self.W.device
        self.maps = {n: torch.empty(self.batch, 1, self.H, self.Wd, device=dev, dtype=torch.float32) for n in self.want}

    def run(self, x: torch.Tensor, want=MAP_NAMES, use_graph: bool = False) -> Dict[str, torch.Tensor]:
        """One forward.  Fresh output tensors per call unless `use_graph` (then the plan's static buffers are
        returned and overwritten by the next call)."""
        self.x_in.copy_(x, non_blocking=True)
        if use_graph:
            if self.graph is None or tuple(want) != tuple(self.want):
                self.want = tuple(want)
                self._alloc_maps()
                self._launch_all()  # warm-up outside capture
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch_all()
                self.graph = g
            self.graph.replay()
            return dict(self.maps)
        self.want = tuple(want)
        self.graph = None
        self._alloc_maps()
        self._launch_all()
        return dict(self.maps)

    def _launch_all(self):
        for s in self.steps:
            s()

    def time_launches(self, iters: int = 3):
        """Per-launch device time with CUDA events on the current stream (a separate, instrumented pass: the
        events serialise nothing but add a little launch overhead, so use it for SHARES and per-kernel rates,
        not for the step time).  Returns [(name, kind, mean ms, algorithmic flops)]."""
        if self.maps is None:
            self._alloc_maps()
        n = len(self.steps)
        acc = [0.0] * n
        for _ in range(iters):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
            ev[0].record()
            for i, s in enumerate(self.steps):
                s()
                ev[i + 1].record()
            torch.cuda.synchronize()
            for i in range(n):
                acc[i] += ev[i].elapsed_time(ev[i + 1])
        return [(nm, self.step_kind.get(nm, "elementwise"), acc[i] / iters, self.step_flops.get(nm, 0.0))
                for i, nm in enumerate(self.names)]

    def launch_bytes(self) -> Dict[str, float]:
        """Algorithmic HBM bytes per igemm launch (for the bandwidth side of the roofline)."""
        return dict(self.step_bytes)

    def launch_write_bytes(self) -> Dict[str, float]:
        """The bytes of launch_bytes() that are written."""
        return dict(self.step_write_bytes)

    def tap(self, name: str) -> torch.Tensor:
        """NCHW fp32 copy of a recorded stage tensor (tests only)."""
        t, c = self.taps[name]
        return t[..., :c].float().permute(0, 3, 1, 2).contiguous()
