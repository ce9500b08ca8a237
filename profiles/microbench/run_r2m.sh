set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 600 python -m pytest tests/test_gpu_lz4.py -x -q 2>&1 | tail -3
timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tail -2
timeout 300 python profiles/microbench/degenerate_lz4.py 2>&1 | tee gpurun_out/lz4_degenerate_r2m.txt
