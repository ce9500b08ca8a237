set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 600 python -m pytest tests/test_gpu_lz4.py tests/test_gpu_inverse_chain.py tests/test_gpu_dropin.py tests/test_gpu_compat.py tests/test_gpu_batch.py -x -q 2>&1 | tail -1
timeout 300 python profiles/microbench/quick_lz4_decode.py 268435456 2>&1 | tail -3
