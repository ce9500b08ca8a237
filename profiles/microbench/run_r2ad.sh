# last run of round 2 with the tree as committed: all GPU tests, smoke, the bench pair
set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2ad.json 2> gpurun_out/bench_r2ad.err; echo bench rc=$?
timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r2ad_reference.json 2>> gpurun_out/bench_r2ad.err; echo ref rc=$?
tail -c 300 gpurun_out/bench_r2ad.err; head -c 300 gpurun_out/bench_r2ad.json; echo
