"""Drop-in SVTRv2 recogniser, inference only (reference: model/rec2/svtrv2.py:409-569).

Same constructor (`variant`, `in_channels`, `charset`, `dropout`, `context_window`), the same 318 `state_dict` keys
(strictly loadable from a reference checkpoint, src/pipeline/pipeline2.py:72-79 -- the training-only SGM parameters are
held but never run), `forward(x) -> log_probs (T, B, num_classes)`, `decode_probs`, `decode_greedy`, `.tokenizer`,
`.blank_id`.  The forward is a static launch list on one B200 (activations = NHWC bf16 token maps):

  * every nn.Linear / nn.Conv2d except the 3-channel stem conv is one tcgen05 implicit-GEMM launch of igemm.cu (Linear =
    1x1 conv over the token image; LocalMixing's grouped 3x3 convs = a 3x3 conv with block-diagonal weights, BatchNorm and
    bias folded into the epilogue; residual adds in the epilogue);
  * GELU runs in the GEMM epilogue (B200_EPI_GELU, before the residual add); LayerNorm, multi-head attention (head_dim 32), the FRM selecting-token attention, PatchMerging's
    row stride and log_softmax are the kernels of svtr.cu; the greedy decode is ctc.cu.

No CPU fallback, no eager torch path: CPU tensors, training mode or `targets` raise.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from ... import ops
from .decode import GreedyCTCDecoder
from .tokenizer import Tokenizer
from .vocab import VOCAB

VARIANTS = {  # reference svtrv2.py:388-404
    "tiny": {"dims": [64, 128, 256], "num_blocks": [3, 6, 3], "num_local": [3, 3, 0]},
    "small": {"dims": [96, 192, 256], "num_blocks": [3, 6, 6], "num_local": [3, 3, 0]},
    "base": {"dims": [128, 256, 384], "num_blocks": [3, 6, 6], "num_local": [3, 2, 0]},
}


# ---------------------------------------------------------------------------------------- parameter holders (no forward)
class _MLP(nn.Module):
    def __init__(self, dim, expansion=4):
        super().__init__()
        self.fc1 = nn.Linear(dim, int(dim * expansion))
        self.fc2 = nn.Linear(int(dim * expansion), dim)


class _Local(nn.Module):
    def __init__(self, dim):
        super().__init__()
        g = max(dim // 32, 1)
        self.conv1 = nn.Conv2d(dim, dim, 3, 1, 1, groups=g)
        self.bn1 = nn.BatchNorm2d(dim)
        self.conv2 = nn.Conv2d(dim, dim, 3, 1, 1, groups=g)
        self.bn2 = nn.BatchNorm2d(dim)


class _Global(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)


class _Block(nn.Module):
    def __init__(self, dim, is_local):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.mixer = _Local(dim) if is_local else _Global(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _MLP(dim)


class _Stage(nn.Module):
    def __init__(self, dim, num_blocks, num_local):
        super().__init__()
        self.blocks = nn.ModuleList(_Block(dim, i < num_local) for i in range(num_blocks))


class _Stem(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout // 2, 3, stride=2, padding=1)
        self.bn1 = nn.BatchNorm2d(cout // 2)
        self.conv2 = nn.Conv2d(cout // 2, cout, 3, stride=2, padding=1)
        self.bn2 = nn.BatchNorm2d(cout)


class _Merge(nn.Module):
    def __init__(self, din, dout):
        super().__init__()
        self.conv = nn.Conv2d(din, dout, 3, stride=(2, 1), padding=1)
        self.norm = nn.BatchNorm2d(dout)


class _FRM(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.h_qkv = nn.Linear(dim, dim * 3)
        self.h_proj = nn.Linear(dim, dim)
        self.h_norm = nn.LayerNorm(dim)
        self.h_mlp = _MLP(dim)
        self.h_norm2 = nn.LayerNorm(dim)
        self.select_token = nn.Parameter(torch.randn(1, 1, dim) * 0.02)
        self.v_q = nn.Linear(dim, dim)
        self.v_kv = nn.Linear(dim, dim * 2)
        self.v_proj = nn.Linear(dim, dim)
        self.v_norm_q = nn.LayerNorm(dim)
        self.v_norm_kv = nn.LayerNorm(dim)
        self.v_mlp = _MLP(dim)
        self.v_norm2 = nn.LayerNorm(dim)


class _SGM(nn.Module):
    """Training-only branch (reference svtrv2.py:253-383): parameters kept for strict state_dict loading, never run."""

    def __init__(self, dim, num_classes):
        super().__init__()
        heads = max(dim // 32, 1)
        self.char_embed = nn.Embedding(num_classes, dim)
        self.context_norm = nn.LayerNorm(dim)
        self.context_attn = nn.MultiheadAttention(dim, heads, batch_first=True)
        self.context_mlp = _MLP(dim, expansion=2)
        self.context_norm2 = nn.LayerNorm(dim)
        self.left_token = nn.Parameter(torch.randn(1, 1, dim) * 0.02)
        self.right_token = nn.Parameter(torch.randn(1, 1, dim) * 0.02)
        self.cross_q = nn.Linear(dim, dim)
        self.cross_kv = nn.Linear(dim, dim * 2)
        self.cross_proj = nn.Linear(dim, dim)
        self.cross_norm_q = nn.LayerNorm(dim)
        self.cross_norm_kv = nn.LayerNorm(dim)
        self.sgm_head = nn.Linear(dim, num_classes)


def _p64(c: int) -> int:
    """Channel counts are padded to whole 64-channel K blocks / 64-column store chunks ('small': 96 -> 128, stem 48 -> 64;
    'tiny': stem 32 -> 64).  Padding weights, biases and activations are zero, LayerNorm runs over the real width."""
    return (c + 63) // 64 * 64


def _pad_w(w: torch.Tensor) -> torch.Tensor:
    o, i = w.shape[0], w.shape[1]
    if o == _p64(o) and i == _p64(i):
        return w
    out = torch.zeros(_p64(o), _p64(i), *w.shape[2:], device=w.device, dtype=w.dtype)
    out[:o, :i] = w
    return out


# ---------------------------------------------------------------------------------------- weight preparation
class _Prepared:
    """bf16 K-major matrices + fp32 (scale, shift) per GEMM, fp32 LayerNorm affines, the constant FRM query."""

    def __init__(self, model: "SVTRv2", device):
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32)
        self.device = device
        self.G: Dict[str, tuple] = {}
        self.LN: Dict[str, tuple] = {}

        def linear(name, lin, extra_shift=None):
            w = ops.pack_conv_weight(_pad_w(f32(lin.weight).reshape(lin.out_features, lin.in_features, 1, 1)))
            shift = f32(lin.bias) if extra_shift is None else f32(lin.bias) + extra_shift
            self.G[name] = (w, torch.ones(w.shape[0], device=device), ops.pad_vec(shift, w.shape[0]))

        def conv_bn(name, conv, bn):
            w = f32(conv.weight)
            if conv.groups > 1:  # grouped -> block-diagonal dense (o, i, 3, 3)
                o, ig = w.shape[0], w.shape[1]
                dense = torch.zeros(o, ig * conv.groups, 3, 3, device=device)
                og = o // conv.groups
                for g in range(conv.groups):
                    dense[g * og:(g + 1) * og, g * ig:(g + 1) * ig] = w[g * og:(g + 1) * og]
                w = dense
            wp = ops.pack_conv_weight(_pad_w(w))
            s, t = ops.fold_bn(f32(bn.weight), f32(bn.bias), f32(bn.running_mean), f32(bn.running_var), bn.eps, f32(conv.bias))
            self.G[name] = (wp, ops.pad_vec(s, wp.shape[0], 1.0), ops.pad_vec(t, wp.shape[0]))

        def ln(name, m):
            self.LN[name] = (f32(m.weight).contiguous(), f32(m.bias).contiguous(), float(m.eps))

        st = model.stem
        s, t = ops.fold_bn(f32(st.bn1.weight), f32(st.bn1.bias), f32(st.bn1.running_mean), f32(st.bn1.running_var), st.bn1.eps,
                           f32(st.conv1.bias))
        mid = st.conv1.out_channels
        w1 = torch.zeros(_p64(mid), 27, device=device)
        w1[:mid] = f32(st.conv1.weight).reshape(mid, 27)
        self.stem1 = (w1, ops.pad_vec(s, _p64(mid), 1.0), ops.pad_vec(t, _p64(mid)))
        conv_bn("stem.conv2", st.conv2, st.bn2)
        for si, stage in enumerate(model.stages):
            for bi, blk in enumerate(stage.blocks):
                p = f"stages.{si}.blocks.{bi}"
                ln(p + ".norm1", blk.norm1)
                ln(p + ".norm2", blk.norm2)
                if isinstance(blk.mixer, _Local):
                    conv_bn(p + ".conv1", blk.mixer.conv1, blk.mixer.bn1)
                    conv_bn(p + ".conv2", blk.mixer.conv2, blk.mixer.bn2)
                else:
                    linear(p + ".qkv", blk.mixer.qkv)
                    linear(p + ".proj", blk.mixer.proj)
                linear(p + ".fc1", blk.mlp.fc1)
                linear(p + ".fc2", blk.mlp.fc2)
        for mi, mg in enumerate(model.merges):
            conv_bn(f"merges.{mi}", mg.conv, mg.norm)
        ln("backbone_norm", model.backbone_norm)
        fr = model.frm
        for n in ("h_norm", "h_norm2", "v_norm_kv", "v_norm2"):
            ln("frm." + n, getattr(fr, n))
        linear("frm.h_qkv", fr.h_qkv)
        linear("frm.h_proj", fr.h_proj)
        linear("frm.h_fc1", fr.h_mlp.fc1)
        linear("frm.h_fc2", fr.h_mlp.fc2)
        linear("frm.v_kv", fr.v_kv)
        tok = f32(fr.select_token).reshape(-1)
        linear("frm.v_proj", fr.v_proj, extra_shift=tok)  # t_q = select_token + v_proj(out)  (svtrv2.py:244)
        linear("frm.v_fc1", fr.v_mlp.fc1)
        linear("frm.v_fc2", fr.v_mlp.fc2)
        # the query of the vertical attention depends on parameters only (svtrv2.py:226-233): v_q(v_norm_q(select_token))
        qn = torch.nn.functional.layer_norm(tok, (tok.numel(),), f32(fr.v_norm_q.weight), f32(fr.v_norm_q.bias), fr.v_norm_q.eps)
        self.frm_q = (f32(fr.v_q.weight) @ qn + f32(fr.v_q.bias)).contiguous()
        linear("head", model.head)
        self.num_classes = model.head.out_features


class _Plan:
    """Static launch list + activation buffers for one batch size (crops are (3, H, W) with H % 8 == 0, W % 4 == 0)."""

    def __init__(self, W: _Prepared, cfg, batch: int, H: int, Wd: int):
        dev = W.device
        self.batch = batch
        import math
        m = 128 // math.gcd(Wd // 4, 128)  # whole 128-token tiles for every token-matrix GEMM, down to the (b * W/4)-row
        b = -(-batch // m) * m             # head: pad the batch (2 for the pipeline's W = 256); padding crops are zeros
        self.bp = b
        real = cfg["dims"]
        dims = [_p64(d) for d in real]  # buffer / GEMM widths; LayerNorm statistics run over real[si] channels
        bf = lambda *s: torch.zeros(*s, device=dev, dtype=torch.bfloat16)  # zero: padding channels are never written
        self.x_in = torch.zeros(b, 3, H, Wd, device=dev, dtype=torch.float32)
        self.steps: List = []
        self.names: List[str] = []
        self.flops = 0.0
        self.keep: List = []

        def add(name, fn):
            self.steps.append(fn)
            self.names.append(name)

        def gemm(name, key, x, y, *, rows=None, geom=None, in_c, out_c, ksize=1, stride=1, stride_w=0, flags=0, residual=None):
            """rows: token-matrix Linear (image = 1 x rows/64 x 64); geom = (batch, h, w): a real convolution."""
            w, s, t = W.G[key]
            if geom is None:
                assert rows % 128 == 0
                geom = (1, rows // 64, 64)  # any factorisation works for a 1x1 conv; 2 rows x 64 tokens = one 128-row tile
            op = ops.Igemm(ops.IGEMM_CONV, x, w, s, t, y, batch=geom[0], in_h=geom[1], in_w=geom[2], in_c=in_c, out_c=out_c,
                           ksize=ksize, stride=stride, stride_w=stride_w, flags=flags, residual=residual,
                           res_ld=out_c if residual is not None else 0, name=name)
            self.flops += op.flops
            add(name, op)

        def lnorm(name, key, x, y, rows, d):
            g, be, eps = W.LN[key]
            add(name, lambda: ops.layernorm(x, y, g, be, rows, g.numel(), eps, x_ld=d, y_ld=d))

        h, w = H // 4, Wd // 4
        mid = _p64(real[0] // 2)
        s1 = bf(b, H // 2, Wd // 2, mid)
        add("stem.conv1", lambda: ops.svtr_stem1(self.x_in, *W.stem1, s1))
        x = bf(b, h, w, dims[0])
        gemm("stem.conv2", "stem.conv2", s1, x, geom=(b, H // 2, Wd // 2), in_c=mid, out_c=dims[0], ksize=3, stride=2,
             flags=ops.EPI_GELU)
        self.taps = {"stem": (x, len(self.steps), real[0])}  # (buffer, steps after which it holds the tap, real width)
        for si in range(3):
            d = dims[si]
            hidw, qkvw = _p64(4 * real[si]), _p64(3 * real[si])  # == 4 * d, 3 * d unless the stage width itself is padded
            rows = b * h * w
            other = bf(b, h, w, d)
            nbuf, cbuf = bf(b, h, w, d), bf(b, h, w, d)
            hid = bf(b, h, w, hidw)
            qkv = bf(b, h, w, qkvw) if cfg["num_local"][si] < cfg["num_blocks"][si] else None
            for bi in range(cfg["num_blocks"][si]):
                p = f"stages.{si}.blocks.{bi}"
                lnorm(p + ".norm1", p + ".norm1", x, nbuf, rows, d)
                if bi < cfg["num_local"][si]:
                    gemm(p + ".conv1", p + ".conv1", nbuf, cbuf, geom=(b, h, w), in_c=d, out_c=d, ksize=3, flags=ops.EPI_GELU)
                    gemm(p + ".conv2", p + ".conv2", cbuf, other, geom=(b, h, w), in_c=d, out_c=d, ksize=3,
                         flags=ops.EPI_GELU | ops.EPI_RES, residual=x)  # x + gelu(bn2(conv2(.))): GELU before the add
                else:
                    assert qkvw == 3 * d, "attention needs an unpadded stage width"
                    gemm(p + ".qkv", p + ".qkv", nbuf, qkv, rows=rows, in_c=d, out_c=3 * d)
                    add(p + ".attn", lambda q=qkv, o=cbuf, n=h * w, d=d: ops.mhsa(q, o, b, n, d))
                    gemm(p + ".proj", p + ".proj", cbuf, other, rows=rows, in_c=d, out_c=d, flags=ops.EPI_RES, residual=x)
                x, other = other, x
                lnorm(p + ".norm2", p + ".norm2", x, nbuf, rows, d)
                gemm(p + ".fc1", p + ".fc1", nbuf, hid, rows=rows, in_c=d, out_c=hidw, flags=ops.EPI_GELU)
                gemm(p + ".fc2", p + ".fc2", hid, other, rows=rows, in_c=hidw, out_c=d, flags=ops.EPI_RES, residual=x)
                x, other = other, x
            self.taps[f"stage{si}"] = (x, len(self.steps), real[si])
            if si < 2:  # PatchMerging: 3x3 conv with stride (2, 1) + BN (two row-parity TMA views, full width)
                dn = dims[si + 1]
                nx = bf(b, h // 2, w, dn)
                gemm(f"merges.{si}", f"merges.{si}", x, nx, geom=(b, h, w), in_c=d, out_c=dn, ksize=3, stride=2, stride_w=1)
                x, h = nx, h // 2
        d = dims[2]
        rows = b * h * w
        xn = bf(b, h, w, d)
        lnorm("backbone_norm", "backbone_norm", x, xn, rows, d)
        # ---- FRM, horizontal: self-attention inside each row of w tokens
        n1, qkv, att, r2, r3 = bf(rows, d), bf(rows, 3 * d), bf(rows, d), bf(rows, d), bf(rows, d)
        hid = bf(rows, 4 * d)
        lnorm("frm.h_norm", "frm.h_norm", xn, n1, rows, d)
        gemm("frm.h_qkv", "frm.h_qkv", n1, qkv, rows=rows, in_c=d, out_c=3 * d)
        add("frm.h_attn", lambda: ops.mhsa(qkv, att, b * h, w, d))
        gemm("frm.h_proj", "frm.h_proj", att, r2, rows=rows, in_c=d, out_c=d, flags=ops.EPI_RES, residual=xn)
        lnorm("frm.h_norm2", "frm.h_norm2", r2, n1, rows, d)
        gemm("frm.h_fc1", "frm.h_fc1", n1, hid, rows=rows, in_c=d, out_c=4 * d, flags=ops.EPI_GELU)
        gemm("frm.h_fc2", "frm.h_fc2", hid, r3, rows=rows, in_c=4 * d, out_c=d, flags=ops.EPI_RES, residual=r2)
        self.taps["frm_h"] = (r3, len(self.steps), real[2])
        # ---- FRM, vertical: the selecting token attends over the h rows of each column
        kv = bf(rows, 2 * d)
        lnorm("frm.v_norm_kv", "frm.v_norm_kv", r3, n1, rows, d)
        gemm("frm.v_kv", "frm.v_kv", n1, kv, rows=rows, in_c=d, out_c=2 * d)
        trows = b * w
        sel, t1, t1n, t2 = bf(trows, d), bf(trows, d), bf(trows, d), bf(trows, d)
        thid = bf(trows, 4 * d)
        add("frm.v_attn", lambda: ops.frm_select(W.frm_q, kv, sel, b, h, w, d))
        gemm("frm.v_proj", "frm.v_proj", sel, t1, rows=trows, in_c=d, out_c=d)
        lnorm("frm.v_norm2", "frm.v_norm2", t1, t1n, trows, d)
        gemm("frm.v_fc1", "frm.v_fc1", t1n, thid, rows=trows, in_c=d, out_c=4 * d, flags=ops.EPI_GELU)
        gemm("frm.v_fc2", "frm.v_fc2", thid, t2, rows=trows, in_c=4 * d, out_c=d, flags=ops.EPI_RES, residual=t1)
        self.taps["frm"] = (t2, len(self.steps), real[2])
        nc = W.num_classes
        self.ld = (nc + 7) // 8 * 8
        self.logits = torch.empty(trows, self.ld, device=dev, dtype=torch.float32)
        op = ops.Igemm(ops.IGEMM_CONV, t2, *W.G["head"], self.logits, batch=1, in_h=trows // 64, in_w=64, in_c=d, out_c=self.ld,
                       ksize=1, stride=1, flags=ops.EPI_OUT_F32, name="head")
        self.flops += op.flops
        add("head", op)
        self.T, self.nc = w, nc
        self.log_probs = torch.empty(w, b, nc, device=dev, dtype=torch.float32)
        add("log_softmax", lambda: ops.log_softmax_tbc(self.logits, self.log_probs, b, w, nc, self.ld))
        self.n_launches = len(self.steps)
        self.graph = None

    def run(self, x: torch.Tensor, taps: Optional[dict] = None, use_graph: bool = False) -> torch.Tensor:
        self.x_in[:self.batch].copy_(x)
        if use_graph and taps is None:  # replay the captured launch list (the result aliases a plan-owned buffer)
            if self.graph is None:
                for s in self.steps:  # warm-up outside capture (lazy function attributes, module loading)
                    s()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for s in self.steps:
                        s()
                self.graph = g
            self.graph.replay()
            return self.log_probs[:, :self.batch]
        for i, s in enumerate(self.steps):
            s()
            if taps is not None:  # buffers are recycled by later blocks: copy a tap right after the step that completes it
                for k, (buf, after, width) in self.taps.items():
                    if after == i + 1:
                        taps[k] = buf.float().reshape(self.bp, -1, buf.shape[-1])[:self.batch, :, :width].clone()
        return self.log_probs[:, :self.batch]


# ---------------------------------------------------------------------------------------- the module
class SVTRv2(nn.Module):
    def __init__(self, variant="small", in_channels=3, charset=VOCAB, dropout=0.0, context_window=3):
        super().__init__()
        assert variant in VARIANTS, f"Unknown variant: {variant}. Choose from {list(VARIANTS.keys())}"
        cfg = VARIANTS[variant]
        dims = cfg["dims"]
        if in_channels != 3:
            raise NotImplementedError("B200 SVTRv2 path: RGB crops only (in_channels == 3, pipeline2.py:72) -- no "
                                      "unaccelerated fallback")
        self.variant, self.cfg = variant, cfg
        self.tokenizer = Tokenizer(charset)
        self.dims = dims
        self.stem = _Stem(in_channels, dims[0])
        self.stages = nn.ModuleList(_Stage(dims[i], cfg["num_blocks"][i], cfg["num_local"][i]) for i in range(3))
        self.merges = nn.ModuleList(_Merge(dims[i], dims[i + 1]) for i in range(2))
        self.backbone_norm = nn.LayerNorm(dims[2])
        self.frm = _FRM(dims[2])
        self.sgm = _SGM(dims[2], self.tokenizer.num_classes)
        self.head = nn.Linear(dims[2], self.tokenizer.num_classes)
        self.blank_id = self.tokenizer.blank_id
        self._prepared: Optional[_Prepared] = None
        self._sig = None
        self._plans: Dict[tuple, _Plan] = {}
        self._decoder = GreedyCTCDecoder(self.tokenizer)
        self.use_cuda_graph = False  # replay a captured graph: log_probs then alias a plan-owned buffer (next call overwrites)

    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _plan(self, x) -> _Plan:
        sig = (x.device, self._signature())
        if self._prepared is None or self._sig != sig:
            self._prepared, self._sig = _Prepared(self, x.device), sig
            self._plans.clear()
        key = tuple(x.shape)
        if key not in self._plans:
            self._plans[key] = _Plan(self._prepared, self.cfg, x.shape[0], x.shape[2], x.shape[3])
        return self._plans[key]

    @torch.no_grad()
    def forward(self, x, targets=None):
        return self.run(x, targets, use_graph=self.use_cuda_graph)

    @torch.no_grad()
    def run(self, x, targets=None, use_graph: bool = False):
        """forward() with the graph option passed per call (pipeline.OCRPipeline uses it; the module's own
        use_cuda_graph attribute is what a plain model(x) follows)."""
        if targets is not None or self.training:
            raise RuntimeError("B200 SVTRv2 path is inference-only: call model.eval() and pass no targets "
                               "(the SGM branch and training, src/rec2/train.py, are out of scope)")
        if not (isinstance(x, torch.Tensor) and x.dim() == 4 and x.shape[1] == 3 and x.is_floating_point()):
            raise ValueError("SVTRv2.forward expects a float (B, 3, H, W) tensor")
        if not x.is_cuda:
            raise RuntimeError("B200 SVTRv2 path runs on CUDA only (there is no CPU fallback)")
        if x.shape[2] % 32 or x.shape[3] % 128 or x.shape[2] > 256 or x.shape[3] > 2048:
            raise ValueError("B200 SVTRv2 path: crops must be (3, H, W) with H a multiple of 32 and W a multiple of 128 "
                             "(the pipeline's 32 x 256): whole 128-token tiles at every stage")
        self._check_attention_lengths(x.shape[2], x.shape[3])
        with torch.cuda.device(x.device):
            lp = self._plan(x).run(x.float().contiguous(), use_graph=use_graph)
            return lp if use_graph else lp.clone()

    MAX_ATTENTION_TOKENS = 768  # b200_mhsa keeps K / V of one (sequence, head) in shared memory

    def _check_attention_lengths(self, H: int, W: int) -> None:
        """Reject crop sizes whose attention sequences b200_mhsa cannot hold BEFORE any plan / buffer is built
        (stage i attends over its whole (H / 4 / 2^i) x (W / 4) token map in the blocks after the first num_local)."""
        for si in range(3):
            n = (H // (4 << si)) * (W // 4)
            if self.cfg["num_blocks"][si] > self.cfg["num_local"][si] and n > self.MAX_ATTENTION_TOKENS:
                raise ValueError(f"B200 SVTRv2 path: a {H} x {W} crop gives stage {si} global attention over {n} tokens; "
                                 f"the attention kernel holds at most {self.MAX_ATTENTION_TOKENS} (e.g. 32 x 768 or 64 x 384)")
        if W // 4 > self.MAX_ATTENTION_TOKENS:
            raise ValueError(f"B200 SVTRv2 path: the FRM's row attention over {W // 4} tokens exceeds {self.MAX_ATTENTION_TOKENS}")

    @torch.no_grad()
    def forward_with_taps(self, x):
        """forward + (B, N, D) fp32 copies of the stage outputs, for the parity tests."""
        self._check_attention_lengths(x.shape[2], x.shape[3])
        with torch.cuda.device(x.device):
            taps = {}
            lp = self._plan(x).run(x.float().contiguous(), taps).clone()
            return lp, taps

    def decode_probs(self, log_probs):
        return self._decoder(log_probs)

    @torch.inference_mode()
    def decode_greedy(self, images):
        self.eval()
        return self.decode_probs(self.forward(images))
