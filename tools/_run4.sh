export B200_DEBUG=16 B200_VERBOSE=1
EPI_SET=stem timeout 120 python tools/epi_bench.py 2>&1 | grep -v "^dcn2"
EPI_SET=offset timeout 120 python tools/epi_bench.py 2>&1 | grep -v "^dcn2\|plan" | head -4
