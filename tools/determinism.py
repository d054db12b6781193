"""Development aid: run the batch-16 forward repeatedly and report the first plan step whose output differs between runs
(every kernel here is meant to be run-to-run deterministic: static tile schedules, no atomics on data)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import randomize_
from ocr_vi_invoice_b200 import ops
from ocr_vi_invoice_b200.model.det.dbnet import DBNetPP
from ocr_vi_invoice_b200.synth import synthetic_batch


def checksum(t):
    v = t.contiguous().view(torch.int16) if t.element_size() == 2 else t.contiguous().view(torch.int32)
    return int(v.to(torch.int64).sum().item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=20)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--test-model", action="store_true")
    ap.add_argument("--poison", action="store_true")
    ap.add_argument("--cycles", type=int, default=0)
    ap.add_argument("--burst", type=int, default=3)
    ap.add_argument("--pack", type=int, default=0, help="GiB block to pre-cache so plan buffers are carved back to back")
    ap.add_argument("--heat", type=int, default=0, help="8192^3 bf16 matmuls between cycles")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    if a.test_model:  # the weights / input of tests/test_gpu_model.py::test_config2_full_size_properties
        from oracle.weights import make_state_dict
        model = DBNetPP(pretrained=False)
        model.load_state_dict(make_state_dict(9, offset_px=1.0), strict=True)
        model = model.to(dev).eval()
        x = torch.from_numpy(synthetic_batch(a.batch, a.size, a.size, seed=4)).to(dev)
    else:
        model = randomize_(DBNetPP(pretrained=False), 0).to(dev).eval()
        x = torch.from_numpy(synthetic_batch(min(a.batch, 4), a.size, a.size, seed=0)).repeat((a.batch + 3) // 4, 1, 1, 1)[:a.batch].to(dev)
    if a.pack:
        # make the caching allocator carve every plan buffer out of ONE cached block: buffers then sit back to back with no
        # slack between them (as inside a long-lived process), so an out-of-bounds access of a kernel hits a live neighbour
        big = torch.empty(a.pack << 30, dtype=torch.uint8, device=dev)
        big.fill_(0x7f)
        del big
    if a.cycles:
        # the sequence of tests/test_gpu_model.py::test_config2_full_size_properties, over and over on fresh plans: first eager
        # forward vs CUDA-graph replays, with per-step checksums to localise a difference; optional heating in between
        for c in range(a.cycles):
            if a.heat:
                hm = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
                for _ in range(a.heat):
                    hm = (hm @ hm).clamp_(-1, 1)
                del hm
            model._plans.clear()
            model.use_cuda_graph = False
            plan = model._get_plan(x)
            outs0 = [(n, s._keep[4]) for n, s in zip(plan.names, plan.steps) if isinstance(s, ops.Igemm)] + \
                    [(n, s._keep[3]) for n, s in zip(plan.names, plan.steps) if isinstance(s, ops.AsfPlan)]
            first = {k: v.clone() for k, v in model(x).items()}
            torch.cuda.synchronize()
            cs0 = {n: checksum(t) for n, t in outs0}
            model.use_cuda_graph = True
            for r in range(3):
                rep = model(x)
                torch.cuda.synchronize()
                cs1 = {n: checksum(t) for n, t in outs0}
                diff = [n for n in plan.names if n in cs0 and cs0[n] != cs1[n]]
                if diff or not torch.equal(first["binary"], rep["binary"]):
                    print(f"cycle {c} replay {r}: differs from the eager run, first step {diff[0] if diff else '?'} ({len(diff)} steps)", flush=True)
            del plan, outs0, first, rep
            model._plans.clear()
            import gc
            gc.collect()
            torch.cuda.empty_cache()
        print(f"{a.cycles} cycles done")
        return
    if a.poison:
        # dirty the caching allocator's blocks first: the plan's torch.empty() buffers then start with NaN / garbage, and the
        # FIRST forward is compared with the second (a kernel that depends on memory it never wrote shows up here)
        for fill in (float("nan"), 1e30, -7.0):
            junk = [torch.full((1 << 28,), fill, device=dev) for _ in range(12)]
            del junk
        model._plans.clear()
        plan = model._get_plan(x)
        outs0 = [(n, s._keep[4]) for n, s in zip(plan.names, plan.steps) if isinstance(s, ops.Igemm)] + \
                [(n, s._keep[3]) for n, s in zip(plan.names, plan.steps) if isinstance(s, ops.AsfPlan)]
        first = {k: v.clone() for k, v in model(x).items()}
        torch.cuda.synchronize()
        cs0 = {n: checksum(t) for n, t in outs0}
        second = model(x)
        torch.cuda.synchronize()
        cs1 = {n: checksum(t) for n, t in outs0}
        diff = [n for n in plan.names if n in cs0 and cs0[n] != cs1[n]]
        print("first vs second forward on poisoned memory:", "identical" if not diff else f"DIFFER from step {diff[0]} ({len(diff)} steps)")
        for k in first:
            if not torch.equal(first[k], second[k]):
                print(f"   map {k} differs in {int((first[k] != second[k]).sum())} elements")
    model(x)
    plan = model._get_plan(x)
    outs = []
    for name, s in zip(plan.names, plan.steps):
        if isinstance(s, ops.Igemm):
            outs.append((name, s._keep[4]))
        elif isinstance(s, ops.AsfPlan):
            outs.append((name, s._keep[3]))
    ref = None
    bad = {}
    for r in range(a.runs):
        plan.x_in.copy_(x)
        plan._launch_all()
        torch.cuda.synchronize()
        cs = [checksum(t) for _, t in outs] + [checksum(plan.maps["binary"])]
        if ref is None:
            ref = cs
            continue
        for (name, _), c0, c1 in zip(outs + [("binary", None)], ref, cs):
            if c0 != c1:
                bad.setdefault(name, 0)
                bad[name] += 1
                print(f"run {r}: first differing step {name}", flush=True)
                break
    print("deterministic" if not bad else f"NON-DETERMINISTIC first-differing steps: {bad}")
    # CUDA-graph replays (kernels back to back, no host gaps) against the eager result
    eager = {k: v.clone() for k, v in model(x).items()}
    model.use_cuda_graph = True
    ndiff = {}
    for r in range(a.runs):
        rep = model(x)
        torch.cuda.synchronize()
        cs = [checksum(t) for _, t in outs]
        for (name, _), c0, c1 in zip(outs, ref, cs):
            if c0 != c1:
                print(f"replay {r}: first differing step {name}", flush=True)
                break
        for k in eager:
            if not torch.equal(rep[k], eager[k]):
                ndiff[k] = ndiff.get(k, 0) + 1
                d = (rep[k] != eager[k])
                idx = d.nonzero()
                print(f"replay {r}: {k} differs in {int(d.sum())} elements, first {idx[0].tolist()} last {idx[-1].tolist()}", flush=True)
                break
    print("graph replays identical to eager" if not ndiff else f"GRAPH REPLAY DIFFERS: {ndiff}")
    # back-to-back forwards with no host synchronisation in between (what a serving loop does): replay, replay, check
    nb = 0
    for r in range(a.runs):
        for _ in range(a.burst):
            rep = model(x)
        torch.cuda.synchronize()
        cs = [checksum(t) for _, t in outs]
        bad_steps = [name for (name, _), c0, c1 in zip(outs, ref, cs) if c0 != c1]
        if bad_steps or not torch.equal(rep["binary"], eager["binary"]):
            nb += 1
            d = (rep["binary"] != eager["binary"])
            imgs = sorted(set(d.nonzero()[:, 0].tolist())) if d.any() else []
            print(f"burst {r}: differs; first step {bad_steps[:3]}; images {imgs}", flush=True)
    print(f"back-to-back replays: {nb} bad bursts of {a.runs}")
    model.use_cuda_graph = False
    nb = 0
    for r in range(a.runs):
        for _ in range(a.burst):
            rep = model(x)
        torch.cuda.synchronize()
        cs = [checksum(t) for _, t in outs]
        bad_steps = [name for (name, _), c0, c1 in zip(outs, ref, cs) if c0 != c1]
        if bad_steps or not torch.equal(rep["binary"], eager["binary"]):
            nb += 1
            print(f"eager burst {r}: differs; first step {bad_steps[:3]}", flush=True)
    print(f"back-to-back eager forwards: {nb} bad bursts of {a.runs}")


if __name__ == "__main__":
    main()
