set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
LJB_BENCH_BATCH_IMAGES=64 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2u.json 2> gpurun_out/bench_r2u.err; echo bench rc=$?
tail -c 400 gpurun_out/bench_r2u.err
