set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_jpeg.py -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2ab.json 2> gpurun_out/bench_r2ab.err; echo bench rc=$?; tail -c 300 gpurun_out/bench_r2ab.err
