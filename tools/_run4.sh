export EPI_SET=conv3 B200_DEBUG=16
timeout 120 python tools/epi_bench.py 2>&1 | grep -v "^dcn2"
