"""SVTRv2-base recogniser on one B200: crops/s and the per-launch table.

    python tools/rec_bench.py [--batch 1024] [--iters 10] [--dump profiles/rXX_rec_launch_table.csv]

Workload: `batch` synthetic 32 x 256 crops (the pipeline's recognition input, pipeline2.py:92-128; BASELINE configs[4]
feeds the crops of 4096 pages through this model), random-init 'base' weights of the reference architecture (checkpoint
absent offline), forward = stem -> 15 mixing blocks -> FRM -> CTC head -> log_softmax, then the greedy decode kernel.
Times are CUDA events on the launching stream; the per-launch pass times every launch alone (L2 flushed before each).
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--dump", default="")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--profile-step", default="", help="comma-separated plan steps to run once between cudaProfilerStart/Stop (ncu --profile-from-start off)")
    args = ap.parse_args()
    from ocr_vi_invoice_b200.model.rec2.svtrv2 import SVTRv2

    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = SVTRv2(variant="base").to(dev).eval()  # random init of the reference architecture (checkpoint absent offline)
    model.use_cuda_graph = not args.no_graph
    x = torch.randn(args.batch, 3, 32, 256, device=dev)
    for _ in range(3):
        lp = model(x)
    torch.cuda.synchronize()
    plan = model._plan(x)
    if args.profile_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()  # ncu --profile-from-start off: only these launches are captured
        for name in args.profile_step.split(","):
            plan.steps[plan.names.index(name)]()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        lp = model(x)
        model._decoder.run_device(lp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    # per-launch table
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for name, step in zip(plan.names, plan.steps):
        t = []
        for _ in range(3):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            b.record()
            b.synchronize()
            t.append(a.elapsed_time(b))
        fl = getattr(step, "flops", 0.0)
        rows.append((name, float(np.median(t)), fl))
    tot = sum(r[1] for r in rows)
    kinds = {}
    for name, t, fl in rows:
        k = ("gemm" if fl else "attn" if "attn" in name else "layernorm" if "norm" in name else "gelu" if "gelu" in name else "other")
        kinds.setdefault(k, [0.0, 0.0])
        kinds[k][0] += t
        kinds[k][1] += fl
    if args.dump:
        with open(args.dump, "w") as f:
            f.write("name,ms,gflop,tflops\n")
            for name, t, fl in rows:
                f.write(f"{name},{t:.4f},{fl / 1e9:.2f},{fl / (t * 1e-3) / 1e12 if t > 0 else 0:.1f}\n")
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    gemm_tf = kinds["gemm"][1] / (kinds["gemm"][0] * 1e-3) / 1e12
    print(json.dumps({
        "metric": "crops/sec", "value": round(args.batch / ms * 1e3, 1), "unit": "crops/s", "batch": args.batch,
        "ms_per_batch": round(ms, 3), "launches": len(rows), "model_gflop_per_crop": round(plan.flops / plan.bp / 1e9, 3),
        "model_tflops": round(plan.flops / (ms * 1e-3) / 1e12, 1),
        "sum_of_launches_ms": round(tot, 3),
        "shares": {k: round(v[0] / tot, 3) for k, v in kinds.items()},
        "gemm": {"achieved": round(gemm_tf, 1), "peak": peak.get("bf16_tflops_sustained"), "unit": "TFLOP/s",
                 "frac": round(gemm_tf / peak["bf16_tflops_sustained"], 3) if peak else None,
                 "note": "dense-equivalent FLOPs: the grouped 3x3 convs run as block-diagonal dense GEMMs"},
        "top": sorted([(n, round(t, 3)) for n, t, _ in rows], key=lambda r: -r[1])[:10]}))


if __name__ == "__main__":
    main()
