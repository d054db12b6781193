export EPI_SET=conv3
echo "== default"; B200_VERBOSE=1 timeout 120 python tools/epi_bench.py 2>&1 | grep -v "^dcn2"
echo "== stages 2"; B200_STAGES=2 timeout 120 python tools/epi_bench.py 2>&1 | grep -v plan
echo "== pair kb>=4"; B200_PAIR_MIN_KB=4 timeout 120 python tools/epi_bench.py 2>&1 | grep -v plan
echo "== pair kb>=4 stages 2"; B200_PAIR_MIN_KB=4 B200_STAGES=2 timeout 120 python tools/epi_bench.py 2>&1 | grep -v plan
unset EPI_SET
timeout 300 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -3
