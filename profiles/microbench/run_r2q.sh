set -x
LJB_LZ4_PHASES=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | grep "ljb lz4" | tail -3
timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tail -1
