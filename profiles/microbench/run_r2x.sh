set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 900 python -m pytest tests/test_gpu_compat.py tests/test_gpu_sharding.py -x -q 2>&1 | tail -3
