timeout 300 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -3
EPI_SET=offset timeout 120 python tools/epi_bench.py 2>&1 | grep " ms " | cut -c1-60
EPI_SET=conv2 timeout 120 python tools/epi_bench.py 2>&1 | grep " ms " | cut -c1-80
timeout 300 python tools/stress_mix.py 300 "offset,conv2,head.0,smooth1" 2>&1 | tail -2
bash tools/ab_bench.sh grp "B200_NO_MMA_GROUP=1"
