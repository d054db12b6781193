set -x
timeout 600 python -m pytest tests/test_gpu_igemm.py -x -q -m gpu -k "conv_igemm" 2>&1 | tail -5
B200_VERBOSE=1 timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu 2>&1 | grep -v "igemm plan" | tail -5
timeout 300 python tools/stress_mix.py 400 "head.0,smooth1,layer1.conv3,head tail" 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/ps_lt.csv 2>gpurun_out/ps.err | cut -c1-300
grep -E "^head.0|^smooth|^lat0" gpurun_out/ps_lt.csv
