"""BASELINE.json configs[2]: modulated DCNv2 layer microbench on the ResNet-50 stage shapes.

For every stage shape (c3 = layer2, c4 = layer3, c5 = layer4; stride 1 and the stride-2 first block) and batch
1..64 it times, on the same GPU and the same inputs (x ~ N(0,1), offsets ~ N(0, 2 px), mask logits ~ N(0,1),
kaiming-uniform(a=1) weights as model/det/dcn.py:37, seed 0):

  * ours:        b200_dcn_coords (raw offset/mask conv output -> sampling records; in the model this is the fused
                 epilogue of the offset conv and costs nothing extra) + the fused sampling/tcgen05 DCN kernel, bf16;
  * torchvision: torchvision.ops.deform_conv2d in fp32 and fp16 (bf16 raises), the reference's call (dcn.py:48).

FLOPs = 2 * b * Ho * Wo * Cout * Cin * 9 (sampling not counted).  Writes one CSV row per (shape, batch) and checks
our output against the torchvision fp32 result (relative rms) on every shape.

    python tools/dcn_microbench.py [--out gpurun_out/dcn_microbench.csv] [--batches 1,2,4,8,16,32,64]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchvision.ops as tvops

from ocr_vi_invoice_b200 import ops

# name, C, input size at batch-16/1024^2 scale, stride
SHAPES = [
    ("c3/layer2 s1", 128, 128, 1), ("c3/layer2 s2", 128, 256, 2),
    ("c4/layer3 s1", 256, 64, 1), ("c4/layer3 s2", 256, 128, 2),
    ("c5/layer4 s1", 512, 32, 1), ("c5/layer4 s2", 512, 64, 2),
]


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/dcn_microbench.csv")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64")
    ap.add_argument("--peak", type=float, default=0.0, help="bf16 TFLOP/s to report fractions against")
    ap.add_argument("--only", default="", help="substring filter on the shape name (e.g. 'layer3 s1'), for ncu captures")
    ap.add_argument("--no-tv", action="store_true", help="skip the torchvision timings (ncu captures)")
    a = ap.parse_args()
    peak = a.peak
    if not peak:
        try:
            peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
        except Exception:
            peak = 1414.8
    dev = "cuda"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    rows = ["shape,cin,cout,in_hw,stride,batch,gflop,ours_dcn_ms,ours_coords_ms,ours_tflops,ours_frac_of_peak,"
            "tv_fp32_ms,tv_fp16_ms,speedup_vs_fp32,speedup_vs_fp16,rel_rms_vs_tv_fp32"]
    for name, c, hw, s in SHAPES:
        if a.only and a.only not in name:
            continue
        for b in [int(v) for v in a.batches.split(",")]:
            g = torch.Generator(device=dev).manual_seed(0)
            x = torch.randn(b, c, hw, hw, device=dev, generator=g)
            wt = torch.empty(c, c, 3, 3, device=dev)
            torch.nn.init.kaiming_uniform_(wt, a=1, generator=g)
            oh = hw // s
            om = torch.zeros(b, oh, oh, 32, device=dev)
            om[..., :18] = torch.randn(b, oh, oh, 18, device=dev, generator=g) * 2.0
            om[..., 18:27] = torch.randn(b, oh, oh, 9, device=dev, generator=g)
            offset = om[..., :18].permute(0, 3, 1, 2).contiguous()
            mask = torch.sigmoid(om[..., 18:27]).permute(0, 3, 1, 2).contiguous()
            xn = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
            wp = ops.pack_conv_weight(wt, ops.choose_n_pad(c))
            n_pad = wp.shape[0]
            sc, sh = torch.ones(n_pad, device=dev), torch.zeros(n_pad, device=dev)
            y = torch.empty(b, oh, oh, c, device=dev, dtype=torch.bfloat16)
            coords = torch.empty(b, oh, oh, 9, 4, device=dev, dtype=torch.int32)
            if ops.dcn_tmem_supported(c):  # c3 / c4: tensor-memory kernel (B200_DCN_V1=1 times the first kernel instead)
                op = ops.Igemm(ops.IGEMM_DCN_TMEM, xn, ops.pack_dcn_weight(wt, n_pad), sc, sh, y, batch=b, in_h=hw, in_w=hw,
                               in_c=c, out_c=c, ksize=3, stride=s, flags=0, offsets=coords, name=name)
            else:
                op = ops.Igemm(ops.IGEMM_DCN, xn, wp, sc, sh, y, batch=b, in_h=hw, in_w=hw, in_c=c, out_c=c, ksize=3,
                               stride=s, flags=0, offsets=coords, name=name)
            iters = max(5, min(50, int(2000 / (b * oh * oh * c * c / 2 ** 24 + 1))))
            t_coords = timed(lambda: ops.dcn_coords(om, coords, b, hw, hw, c, s), iters)
            t_dcn = timed(op, iters)
            xb, wb = x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float()
            ref = tvops.deform_conv2d(xb, offset, wb, None, stride=s, padding=1, mask=mask)
            out = y.permute(0, 3, 1, 2).float()
            rel = ((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
            assert rel < 2e-2, f"{name} b{b}: relative rms {rel}"
            if a.no_tv:
                t32 = t16 = float("nan")
            else:
                t32 = timed(lambda: tvops.deform_conv2d(x, offset, wt, None, stride=s, padding=1, mask=mask), max(3, iters // 5))
                xh, oh16, wh, mh = x.half(), offset.half(), wt.half(), mask.half()
                t16 = timed(lambda: tvops.deform_conv2d(xh, oh16, wh, None, stride=s, padding=1, mask=mh), max(3, iters // 5))
            gflop = 2.0 * b * oh * oh * c * c * 9 / 1e9
            tf = gflop / t_dcn
            rows.append(f"{name},{c},{c},{hw},{s},{b},{gflop:.2f},{t_dcn:.4f},{t_coords:.4f},{tf:.1f},{tf / peak:.3f},"
                        f"{t32:.3f},{t16:.3f},{t32 / (t_dcn + t_coords):.1f},{t16 / (t_dcn + t_coords):.1f},{rel:.4f}")
            print(rows[-1], flush=True)
            del op
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        f.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
