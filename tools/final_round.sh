set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --dump-launches gpurun_out/launch_table_final.csv > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"igemm_kernel|asf_|db_maps|db_head|maxpool|stem_pack|pp_" -s 460 -c 200 --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-200
python tools/rec_bench.py --batch 1024 --dump gpurun_out/rec_launch_table.csv 2>&1 | tail -1 | cut -c1-300
