timeout 300 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -3
bash tools/ab_bench.sh lag "B200_EPI_LAG=0"
