set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 900 python -m pytest tests/test_gpu_lz4.py tests/test_gpu_batch.py tests/test_gpu_errors.py tests/test_gpu_dropin.py tests/test_gpu_compat.py -x -q 2>&1 | tail -3
LJB_PIPE_CHUNK_BYTES=100000 timeout 300 python -m pytest tests/test_gpu_lz4.py -x -q -k "roundtrip or decoder" 2>&1 | tail -2
timeout 200 python profiles/microbench/quick_lz4_decode.py 1073741824 2>&1 | tee gpurun_out/lz4_decode_r2l.txt
