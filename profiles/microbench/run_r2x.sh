set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 600 python -m pytest tests/test_gpu_lz4.py -x -q --tb=short 2>&1 | tail -40
