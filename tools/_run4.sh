export B200_DEBUG=16
EPI_SET=offset timeout 120 python tools/epi_bench.py 2>&1 | grep -v "^dcn2"
EPI_SET=conv2 timeout 120 python tools/epi_bench.py 2>&1 | grep -v "^dcn2"
EPI_SET=conv3 timeout 120 python tools/epi_bench.py 2>&1 | grep -v "^dcn2" | head -10
