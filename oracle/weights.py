"""Seeded, reference-free construction of a DBNetPP(resnet50, dcn=True) state_dict -- TEST INFRASTRUCTURE.

The checkpoint weights/det/best_model.pth is absent (git-ignored upstream), and a random-init reference
model is degenerate for parity purposes: offset_mask_conv is zero-initialised (model/det/dcn.py:28-29) so
deformable sampling never leaves the integer grid, and BN running stats are (0, 1).  `make_state_dict`
therefore draws every tensor from a seeded CPU generator with non-trivial offsets / masks / BN statistics.
The key set (754 keys, 410 unique tensors) follows model/det/backbone.py:34-37 (the torchvision net is
registered twice: `backbone.model.*` and `backbone.layer{1..4}.*`), model/det/neck.py:14-24,55 and
model/det/head.py:11-26.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

LAYERS = ((1, 3, 64), (2, 4, 128), (3, 6, 256), (4, 3, 512))  # (index, blocks, width) of resnet50


def _bn(prefix):
    return [(prefix + ".weight", "bn_w"), (prefix + ".bias", "bn_b"), (prefix + ".running_mean", "bn_m"),
            (prefix + ".running_var", "bn_v"), (prefix + ".num_batches_tracked", "nbt")]


def unique_param_table():
    """[(key under its canonical name, shape, kind)] for the 410 unique tensors, registration order."""
    t = []
    add = lambda k, s, kind: t.append((k, tuple(s), kind))
    add("backbone.model.conv1.weight", (64, 3, 7, 7), "conv")
    for k, kind in _bn("backbone.model.bn1"):
        add(k, () if kind == "nbt" else (64,), kind)
    inplanes = 64
    for li, nblocks, width in LAYERS:
        for b in range(nblocks):
            p = f"backbone.model.layer{li}.{b}"
            out = width * 4
            add(p + ".conv1.weight", (width, inplanes, 1, 1), "conv")
            for k, kind in _bn(p + ".bn1"):
                add(k, () if kind == "nbt" else (width,), kind)
            add(p + ".conv2.weight", (width, width, 3, 3), "conv")
            if li >= 2:
                add(p + ".conv2.offset_mask_conv.weight", (27, width, 3, 3), "offset_w")
                add(p + ".conv2.offset_mask_conv.bias", (27,), "offset_b")
            for k, kind in _bn(p + ".bn2"):
                add(k, () if kind == "nbt" else (width,), kind)
            add(p + ".conv3.weight", (out, width, 1, 1), "conv")
            for k, kind in _bn(p + ".bn3"):
                add(k, () if kind == "nbt" else (out,), kind)
            if b == 0:
                add(p + ".downsample.0.weight", (out, inplanes, 1, 1), "conv")
                for k, kind in _bn(p + ".downsample.1"):
                    add(k, () if kind == "nbt" else (out,), kind)
            inplanes = out
    add("backbone.model.fc.weight", (1000, 2048), "fc")
    add("backbone.model.fc.bias", (1000,), "bias")
    for i, c in enumerate((256, 512, 1024, 2048)):
        add(f"neck.lateral_convs.{i}.weight", (256, c, 1, 1), "conv")
        add(f"neck.lateral_convs.{i}.bias", (256,), "bias")
    for i in range(4):
        add(f"neck.fpn_convs.{i}.conv.weight", (256, 256, 3, 3), "conv")
        for k, kind in _bn(f"neck.fpn_convs.{i}.bn"):
            add(k, () if kind == "nbt" else (256,), kind)
    add("neck.asf.conv_atten.weight", (4, 1024, 1, 1), "atten")
    add("neck.asf.conv_atten.bias", (4,), "bias")
    for br in ("bin_conv", "thresh_conv"):
        add(f"head.{br}.0.conv.weight", (64, 256, 3, 3), "conv")
        for k, kind in _bn(f"head.{br}.0.bn"):
            add(k, () if kind == "nbt" else (64,), kind)
        add(f"head.{br}.1.weight", (64, 64, 2, 2), "convT")
        add(f"head.{br}.1.bias", (64,), "bias")
        for k, kind in _bn(f"head.{br}.2"):
            add(k, () if kind == "nbt" else (64,), kind)
        add(f"head.{br}.4.weight", (64, 1, 2, 2), "convT_last")
        add(f"head.{br}.4.bias", (1,), "bias")
    return t


def alias_of(key: str):
    """`backbone.model.X` -> its second registration under `backbone.layerN.*` (backbone.py:34-37), or None."""
    pre = "backbone.model."
    if not key.startswith(pre):
        return None
    rest = key[len(pre):]
    if rest.startswith("conv1."):
        return "backbone.layer1.0." + rest[len("conv1."):]
    if rest.startswith("bn1."):
        return "backbone.layer1.1." + rest[len("bn1."):]
    if rest.startswith("layer1."):
        return "backbone.layer1.4." + rest[len("layer1."):]
    for n in (2, 3, 4):
        if rest.startswith(f"layer{n}."):
            return f"backbone.layer{n}." + rest[len(f"layer{n}."):]
    return None  # fc.* has a single registration


def state_dict_keys():
    """All 754 keys in the order nn.Module.state_dict() of the reference emits them."""
    table = unique_param_table()
    model_keys = [k for k, _, _ in table if k.startswith("backbone.model.")]
    keys = list(model_keys)
    for n in (1, 2, 3, 4):
        keys += [alias_of(k) for k in model_keys if (alias_of(k) or "").startswith(f"backbone.layer{n}.")]
    keys += [k for k, _, _ in table if not k.startswith("backbone.")]
    return keys


def make_state_dict(seed: int = 0, offset_px: float = 1.5) -> "OrderedDict[str, torch.Tensor]":
    """Seeded state_dict with exercised deformable offsets and non-trivial BN statistics (fp32, CPU)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape, kind in unique_param_table():
        if kind == "nbt":
            v = torch.tensor(100, dtype=torch.long)
        elif kind == "conv":  # He-style so activations keep O(1) scale through the net
            fan_in = shape[1] * shape[2] * shape[3]
            v = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif kind == "convT":
            v = torch.randn(shape, generator=g) * math.sqrt(2.0 / shape[0])
        elif kind == "convT_last":
            v = torch.randn(shape, generator=g) * (2.0 / math.sqrt(shape[0]))
        elif kind == "fc":
            v = torch.randn(shape, generator=g) * 0.01
        elif kind == "atten":
            v = torch.randn(shape, generator=g) * 0.05
        elif kind == "offset_w":  # offsets of a few pixels for O(1) inputs; mask logits O(1)
            fan_in = shape[1] * 9
            v = torch.randn(shape, generator=g) * (offset_px / math.sqrt(fan_in))
            v[18:] *= 1.0 / offset_px
        elif kind == "offset_b":
            v = torch.randn(shape, generator=g) * 0.5
        elif kind == "bias":
            v = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_w":
            v = 0.6 + 0.4 * torch.rand(shape, generator=g)
        elif kind == "bn_b":
            v = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_m":
            v = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_v":
            v = 0.5 + torch.rand(shape, generator=g)
        else:  # pragma: no cover
            raise KeyError(kind)
        sd[key] = v
    out = OrderedDict()
    for k in state_dict_keys():
        if k in sd:
            out[k] = sd[k]
        else:  # alias: share storage with its backbone.model.* twin, as the reference's state_dict does
            pre = {"backbone.layer1.0.": "backbone.model.conv1.", "backbone.layer1.1.": "backbone.model.bn1.",
                   "backbone.layer1.4.": "backbone.model.layer1.", "backbone.layer2.": "backbone.model.layer2.",
                   "backbone.layer3.": "backbone.model.layer3.", "backbone.layer4.": "backbone.model.layer4."}
            for a, m in pre.items():
                if k.startswith(a):
                    out[k] = sd[m + k[len(a):]]
                    break
            else:  # pragma: no cover
                raise KeyError(k)
    return out
