"""CPU restatement of the SVTRv2 recogniser's INFERENCE forward -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

SURVEY.md 8(f) item 3.  Reference (file:line under the reference tree, model/rec2/svtrv2.py):
  ConvStem :106-122, MixingBlock :90-101 (LocalMixing :42-64, GlobalMixing :67-87, MLP :28-39), SVTRStage :143-155,
  PatchMerging :125-138, FRM :160-249, SVTRv2.extract_features / forward :475-544 (inference branch: the SGM of
  :253-383 is training-only and never runs, but its parameters are part of the state_dict).

Functional, state_dict-driven fp32 torch on the CPU; every function cites the lines it follows.  Third-party arithmetic:
torch's conv2d / linear / layer_norm / softmax / gelu (erf form) -- unpinned torch, 2.11 here.
PINNED: tests/golden/svtrv2_small.npz holds the log-probabilities and stage taps of the reference's OWN module
(`SVTRv2(variant='base')` imported from /root/reference by tests/golden/make_svtrv2_golden.py, loaded strictly with
`make_state_dict`); this restatement reproduces them to 1e-4 (tests/test_oracle.py).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

VARIANTS = {  # svtrv2.py:388-404
    "tiny": {"dims": [64, 128, 256], "num_blocks": [3, 6, 3], "num_local": [3, 3, 0]},
    "small": {"dims": [96, 192, 256], "num_blocks": [3, 6, 6], "num_local": [3, 3, 0]},
    "base": {"dims": [128, 256, 384], "num_blocks": [3, 6, 6], "num_local": [3, 2, 0]},
}
BN_EPS = 1e-5
LN_EPS = 1e-5


# ------------------------------------------------------------------------------------------ parameters
def param_shapes(variant: str = "base", num_classes: int = 232, in_channels: int = 3, context_window: int = 3):
    """Ordered {key: shape} of SVTRv2(variant).state_dict() (318 entries for 'base'), built from the architecture."""
    cfg = VARIANTS[variant]
    dims = cfg["dims"]
    out: Dict[str, tuple] = {}

    def bn(p, c):
        out[p + ".weight"] = (c,); out[p + ".bias"] = (c,)
        out[p + ".running_mean"] = (c,); out[p + ".running_var"] = (c,); out[p + ".num_batches_tracked"] = ()

    def ln(p, c):
        out[p + ".weight"] = (c,); out[p + ".bias"] = (c,)

    def lin(p, i, o):
        out[p + ".weight"] = (o, i); out[p + ".bias"] = (o,)

    def conv(p, i, o, groups=1):
        out[p + ".weight"] = (o, i // groups, 3, 3); out[p + ".bias"] = (o,)

    def mlp(p, d, e=4):
        lin(p + ".fc1", d, d * e); lin(p + ".fc2", d * e, d)

    mid = dims[0] // 2
    conv("stem.conv1", in_channels, mid); bn("stem.bn1", mid)
    conv("stem.conv2", mid, dims[0]); bn("stem.bn2", dims[0])
    merges = {}
    for s in range(3):
        d = dims[s]
        for b in range(cfg["num_blocks"][s]):
            p = f"stages.{s}.blocks.{b}"
            ln(p + ".norm1", d)
            if b < cfg["num_local"][s]:
                g = max(d // 32, 1)
                conv(p + ".mixer.conv1", d, d, g); bn(p + ".mixer.bn1", d)
                conv(p + ".mixer.conv2", d, d, g); bn(p + ".mixer.bn2", d)
            else:
                lin(p + ".mixer.qkv", d, 3 * d); lin(p + ".mixer.proj", d, d)
            ln(p + ".norm2", d)
            mlp(p + ".mlp", d)
    # nn.ModuleList order in the reference state_dict: stem, stages, merges, backbone_norm, frm, sgm, head
    for s in range(2):
        conv(f"merges.{s}.conv", dims[s], dims[s + 1]); bn(f"merges.{s}.norm", dims[s + 1])
    d = dims[2]
    ln("backbone_norm", d)
    out["frm.select_token"] = (1, 1, d)
    lin("frm.h_qkv", d, 3 * d); lin("frm.h_proj", d, d); ln("frm.h_norm", d); mlp("frm.h_mlp", d); ln("frm.h_norm2", d)
    lin("frm.v_q", d, d); lin("frm.v_kv", d, 2 * d); lin("frm.v_proj", d, d)
    ln("frm.v_norm_q", d); ln("frm.v_norm_kv", d); mlp("frm.v_mlp", d); ln("frm.v_norm2", d)
    out["sgm.left_token"] = (1, 1, d); out["sgm.right_token"] = (1, 1, d)
    out["sgm.char_embed.weight"] = (num_classes, d)
    ln("sgm.context_norm", d)
    out["sgm.context_attn.in_proj_weight"] = (3 * d, d); out["sgm.context_attn.in_proj_bias"] = (3 * d,)
    lin("sgm.context_attn.out_proj", d, d)
    mlp("sgm.context_mlp", d, 2); ln("sgm.context_norm2", d)
    lin("sgm.cross_q", d, d); lin("sgm.cross_kv", d, 2 * d); lin("sgm.cross_proj", d, d)
    ln("sgm.cross_norm_q", d); ln("sgm.cross_norm_kv", d)
    lin("sgm.sgm_head", d, num_classes)
    lin("head", d, num_classes)
    return out


def make_state_dict(variant: str = "base", seed: int = 0, num_classes: int = 232) -> Dict[str, torch.Tensor]:
    """Seeded, NON-degenerate parameters (the reference's init has zero biases and identity norms): BN statistics and
    affines, LN affines and every bias are randomised so a wrong fold / wrong epsilon shows up in the outputs."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(variant, num_classes).items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(100, dtype=torch.long)
        elif k.endswith("running_var"):
            sd[k] = torch.rand(shp, generator=g) * 1.0 + 0.5
        elif k.endswith("running_mean"):
            sd[k] = torch.randn(shp, generator=g) * 0.2
        elif "norm" in k.split(".")[-2] or ".bn" in k:
            sd[k] = (torch.rand(shp, generator=g) * 0.4 + 0.8) if k.endswith("weight") else torch.randn(shp, generator=g) * 0.1
        elif k.endswith("token"):
            sd[k] = torch.randn(shp, generator=g) * 0.5
        elif k.endswith("bias"):
            sd[k] = torch.randn(shp, generator=g) * 0.1
        elif len(shp) == 4:
            fan_in = shp[1] * 9
            sd[k] = torch.randn(shp, generator=g) * (1.5 / fan_in) ** 0.5
        else:
            sd[k] = torch.randn(shp, generator=g) * (1.5 / shp[-1]) ** 0.5
    return sd


# ------------------------------------------------------------------------------------------ forward
def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], LN_EPS)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _mlp(sd, p, x):  # svtrv2.py:38-39 (nn.GELU() = exact erf form)
    return _lin(sd, p + ".fc2", F.gelu(_lin(sd, p + ".fc1", x)))


def _mhsa(q, k, v, scale):  # svtrv2.py:82-86
    attn = (q @ k.transpose(-2, -1)) * scale
    return attn.softmax(dim=-1) @ v


def local_mixing(sd, p, x, H, W):  # svtrv2.py:57-64
    B, N, D = x.shape
    g = max(D // 32, 1)
    x = x.transpose(1, 2).reshape(B, D, H, W)
    x = F.gelu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], 1, 1, 1, g)))
    x = F.gelu(_bn(sd, p + ".bn2", F.conv2d(x, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], 1, 1, 1, g)))
    return x.flatten(2).transpose(1, 2)


def global_mixing(sd, p, x):  # svtrv2.py:77-87
    B, N, D = x.shape
    heads = max(D // 32, 1)
    hd = D // heads
    qkv = _lin(sd, p + ".qkv", x).reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    out = _mhsa(qkv[0], qkv[1], qkv[2], hd ** -0.5).transpose(1, 2).reshape(B, N, D)
    return _lin(sd, p + ".proj", out)


def frm(sd, x, H, W, taps=None):  # svtrv2.py:196-249
    B, N, D = x.shape
    heads = max(D // 32, 1)
    hd = D // heads
    scale = hd ** -0.5
    rows = x.reshape(B * H, W, D)
    qkv = _lin(sd, "frm.h_qkv", _ln(sd, "frm.h_norm", rows)).reshape(B * H, W, 3, heads, hd).permute(2, 0, 3, 1, 4)
    out = _mhsa(qkv[0], qkv[1], qkv[2], scale).transpose(1, 2).reshape(B * H, W, D)
    rows = rows + _lin(sd, "frm.h_proj", out)
    rows = rows + _mlp(sd, "frm.h_mlp", _ln(sd, "frm.h_norm2", rows))
    if taps is not None:
        taps["frm_h"] = rows.reshape(B, H * W, D)
    cols = rows.reshape(B, H, W, D).permute(0, 2, 1, 3).reshape(B * W, H, D)
    t_q = sd["frm.select_token"].expand(B, W, -1).reshape(B * W, 1, D)
    q = _lin(sd, "frm.v_q", _ln(sd, "frm.v_norm_q", t_q)).reshape(B * W, 1, heads, hd).permute(0, 2, 1, 3)
    kv = _lin(sd, "frm.v_kv", _ln(sd, "frm.v_norm_kv", cols)).reshape(B * W, H, 2, heads, hd).permute(2, 0, 3, 1, 4)
    out = _mhsa(q, kv[0], kv[1], scale).transpose(1, 2).reshape(B * W, 1, D)
    t_q = t_q + _lin(sd, "frm.v_proj", out)
    t_q = t_q + _mlp(sd, "frm.v_mlp", _ln(sd, "frm.v_norm2", t_q))
    return t_q.reshape(B, W, D)


@torch.no_grad()
def svtrv2_forward(sd, x: torch.Tensor, variant: str = "base", taps: Optional[dict] = None) -> torch.Tensor:
    """x: (B, 3, H, W) float32 (32 x 256 crops) -> log_probs (T = W/4, B, num_classes), svtrv2.py:509-544.
    taps (optional dict) receives 'stem', 'stage0..2' (B, N, D token maps), 'frm_h', 'frm' and 'logits'."""
    cfg = VARIANTS[variant]
    x = F.gelu(_bn(sd, "stem.bn1", F.conv2d(x, sd["stem.conv1.weight"], sd["stem.conv1.bias"], 2, 1)))  # :118-122
    x = F.gelu(_bn(sd, "stem.bn2", F.conv2d(x, sd["stem.conv2.weight"], sd["stem.conv2.bias"], 2, 1)))
    B, D, H, W = x.shape
    x = x.flatten(2).transpose(1, 2)
    if taps is not None:
        taps["stem"] = x
    for s in range(3):
        for b in range(cfg["num_blocks"][s]):  # :97-101
            p = f"stages.{s}.blocks.{b}"
            y = _ln(sd, p + ".norm1", x)
            y = local_mixing(sd, p + ".mixer", y, H, W) if b < cfg["num_local"][s] else global_mixing(sd, p + ".mixer", y)
            x = x + y
            x = x + _mlp(sd, p + ".mlp", _ln(sd, p + ".norm2", x))
        if taps is not None:
            taps[f"stage{s}"] = x
        if s < 2:  # PatchMerging :132-138: 3x3 conv, stride (2, 1), then BatchNorm (no activation)
            p = f"merges.{s}"
            y = x.transpose(1, 2).reshape(B, x.shape[2], H, W)
            y = _bn(sd, p + ".norm", F.conv2d(y, sd[p + ".conv.weight"], sd[p + ".conv.bias"], (2, 1), 1))
            H, W = y.shape[2], y.shape[3]
            x = y.flatten(2).transpose(1, 2)
    x = _ln(sd, "backbone_norm", x)
    feats = frm(sd, x, H, W, taps)
    if taps is not None:
        taps["frm"] = feats
    logits = _lin(sd, "head", feats)
    if taps is not None:
        taps["logits"] = logits
    return F.log_softmax(logits.permute(1, 0, 2), dim=-1)
