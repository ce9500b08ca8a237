# last run of round 2 with the tree as committed: all GPU tests, smoke, the bench pair
set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2ae.json 2> gpurun_out/bench_r2ae.err; echo bench rc=$?
timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r2ae_reference.json 2>> gpurun_out/bench_r2ae.err; echo ref rc=$?
tail -c 300 gpurun_out/bench_r2ae.err; head -c 300 gpurun_out/bench_r2ae.json; echo
LJB_LZ4_PHASES=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_phases_r2ae.txt 2>&1; tail -3 gpurun_out/lz4_phases_r2ae.txt
