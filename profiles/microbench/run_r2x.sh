set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 600 python -m pytest tests/test_gpu_lz4.py tests/test_gpu_sharding.py tests/test_gpu_dropin.py tests/test_gpu_compat.py -x -q 2>&1 | tail -1
for n in 4294967296 1073741824 268435456 67108864 16777216; do timeout 300 python profiles/microbench/quick_e2e.py $n 2048 2>&1 | grep "lz4 e2e" | tail -1; done
