# final numbers of round 2 with the kernels as committed: tests, smoke, bench + reference arm, launch list, degenerate / block-length tables
set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2aa.json 2> gpurun_out/bench_r2aa.err; echo bench rc=$?
timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r2aa_reference.json 2>> gpurun_out/bench_r2aa.err; echo ref rc=$?
LJB_BENCH_BATCH_IMAGES=64 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-sample > gpurun_out/plain_launch.log 2>&1 &&
LJB_BENCH_BATCH_IMAGES=64 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_r2aa.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-sample > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_encode -s 1 -c 1 -f -o gpurun_out/lz4_r2aa python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_ncu.log 2>&1
LJB_LZ4_PHASES=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_phases_r2aa.txt 2>&1
timeout 300 python profiles/microbench/degenerate_lz4.py > gpurun_out/lz4_degenerate_r2aa.txt 2>&1
timeout 300 python profiles/microbench/blocklen_lz4.py > gpurun_out/lz4_blocklen_r2aa.txt 2>&1
tail -3 gpurun_out/lz4_phases_r2aa.txt; cat gpurun_out/lz4_degenerate_r2aa.txt gpurun_out/lz4_blocklen_r2aa.txt
tail -c 300 gpurun_out/bench_r2aa.err; head -c 400 gpurun_out/bench_r2aa.json; echo
for c in 4 3; do for s in 0 1; do timeout 300 python profiles/tools/jfif_time.py --dim 16384 --sub $s --comp $c --iters 4 --e2e 2>&1 | tail -2; done; done > gpurun_out/jfif_timing_r2aa.txt
timeout 300 python profiles/microbench/quick_e2e_jpeg.py > gpurun_out/jpeg_e2e_r2aa.txt 2>&1
cat gpurun_out/jfif_timing_r2aa.txt gpurun_out/jpeg_e2e_r2aa.txt
