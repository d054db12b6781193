#!/bin/bash
# end-of-round evidence run (round 2), part A: bench (both arms), launch table, DCN microbench, ncu launch list
# (part B = tools/final_round2_ncu.sh: --set full captures of the kernels of this round, exported as CSV on the box)
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --dump-launches gpurun_out/r02_launch_table.csv > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 400 gpurun_out/r02_bench.json; tail -2 gpurun_out/r02_bench.err
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference.json; cut -c1-260 gpurun_out/r02_bench_reference.json
python tools/dcn_microbench.py --batches 1,4,16,64 --out gpurun_out/r02_dcn_microbench.csv 2>&1 | tail -3
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"igemm_kernel|dcn2_kernel|asf_|db_maps|maxpool|stem_pack|pp_" -s 465 -c 200 --csv --log-file gpurun_out/r02_ncu_launch_list.csv $B > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-200
