timeout 300 python -m pytest tests/test_gpu_igemm.py -x -q -m gpu -k "conv_igemm or chained" 2>&1 | tail -3
bash tools/ab_bench.sh respf "B200_RES_PF=0"
bash tools/ab_bench.sh respf4 "B200_RES_PF=4" | tail -12
bash tools/ab_bench.sh sh256 "B200_SHARE_N256=1" | tail -8
