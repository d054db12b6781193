#!/bin/bash
# development A/B: bench with and without an environment switch, per-launch deltas
#   tools/ab_bench.sh NAME "ENV=VAL"     (B runs with ENV=VAL, A without)
name=$1; envb=$2
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/${name}_lt_a.csv > gpurun_out/${name}_a.json 2> gpurun_out/${name}_a.err; echo rc=$?
env $envb timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/${name}_lt_b.csv > gpurun_out/${name}_b.json 2> gpurun_out/${name}_b.err; echo rc=$?
python - <<PY
import json,csv
for f in ("${name}_a","${name}_b"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "igemm frac", d["roofline"]["frac"], "dcn frac", d["dcn"]["frac"])
    except Exception as e: print(f, "ERR", e)
a={r[0]:float(r[2]) for r in list(csv.reader(open("gpurun_out/${name}_lt_a.csv")))[1:]}
b={r[0]:float(r[2]) for r in list(csv.reader(open("gpurun_out/${name}_lt_b.csv")))[1:]}
for k in a:
    if abs(a[k]-b.get(k,0))>0.003: print(k, "A", a[k], "B", b.get(k))
print("sum A", round(sum(a.values()),3), "sum B", round(sum(b.values()),3))
PY
