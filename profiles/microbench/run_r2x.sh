set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 900 python -m pytest tests/test_gpu_jfif.py tests/test_gpu_dropin.py tests/test_gpu_errors.py tests/test_gpu_compat.py -x -q 2>&1 | tail -3
for c in 4 3; do for s in 0 1; do timeout 300 python profiles/tools/jfif_time.py --dim 16384 --sub $s --comp $c --iters 4 --e2e 2>&1 | tail -2; done; done > gpurun_out/jfif_timing_r2x.txt; cat gpurun_out/jfif_timing_r2x.txt
