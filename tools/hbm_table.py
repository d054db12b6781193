import sys, json, torch
sys.path.insert(0, '.')
from bench import randomize_, peaks
from ocr_vi_invoice_b200.model.det.dbnet import DBNetPP
from ocr_vi_invoice_b200.synth import synthetic_batch
m = randomize_(DBNetPP(pretrained=False), 0).cuda().eval()
x = torch.from_numpy(synthetic_batch(4, 1024, 1024, seed=0)).repeat(4, 1, 1, 1).cuda()
m.use_cuda_graph = False
L = m.time_launches(x, iters=3)
lb = m._get_plan(x).launch_bytes()
pk = peaks(); bal = pk["tf_sustained"] * 1e12 / (pk["hbm_gbs"] * 1e9)
rows = []
for n, kind, ms, fl in L:
    if n in lb:
        rows.append((n, ms, lb[n] / 1e6, lb[n] / ms / 1e6, fl / lb[n], (ms - lb[n] / 6448.4e6) ))
rows.sort(key=lambda r: -r[5])
print("name ms MB GB/s flop/byte excess_ms_over_hbm_floor")
for r in rows[:30]:
    print(f"{r[0]:18s} {r[1]:.3f} {r[2]:8.1f} {r[3]:7.0f} {r[4]:7.1f} {r[5]:.3f}")
print("\nHBM-bound launches (flop/byte below the machine balance %.0f), by time:" % bal)
hb = sorted([r for r in rows if r[4] < bal], key=lambda r: -r[1])
tot_ms = sum(r[1] for r in hb); tot_mb = sum(r[2] for r in hb)
for r in hb:
    print(f"{r[0]:18s} {r[1]:.3f} ms {r[2]:8.1f} MB {r[3]:7.0f} GB/s  flop/byte {r[4]:6.1f}")
print(f"total {tot_ms:.3f} ms, {tot_mb / tot_ms:.0f} GB/s on aggregate")
