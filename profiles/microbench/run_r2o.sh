set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 300 python -m pytest tests/test_gpu_lz4.py -x -q 2>&1 | tail -1
timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tail -2
LJB_LZ4_PHASES=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | grep "ljb lz4" | tail -2
