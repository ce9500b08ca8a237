# round-2 final evidence run: tests, smoke, bench + reference arm, launch list with DRAM traffic, ncu --set full of the hot kernels
set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2v.json 2> gpurun_out/bench_r2v.err; echo bench rc=$?
timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r2v_reference.json 2>> gpurun_out/bench_r2v.err; echo ref rc=$?
LJB_BENCH_BATCH_IMAGES=64 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-sample > gpurun_out/plain_launch.log 2>&1 &&
LJB_BENCH_BATCH_IMAGES=64 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches_r2v.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-sample > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_encode -s 1 -c 1 -f -o gpurun_out/lz4_r2v python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jfif_encode -s 1 -c 1 -f -o gpurun_out/jfif444_r2v python profiles/tools/jfif_time.py --dim 16384 --sub 0 --iters 2 > gpurun_out/jfif_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jfif_encode -s 1 -c 1 -f -o gpurun_out/jfif420_r2v python profiles/tools/jfif_time.py --dim 16384 --sub 1 --iters 2 > gpurun_out/jfif_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jpeg_encode -s 1 -c 1 -f -o gpurun_out/jpeg_r2v python profiles/microbench/quick_jpeg.py 16384 > gpurun_out/jpeg_ncu.log 2>&1
LJB_LZ4_PHASES=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_phases_r2v.txt 2>&1
timeout 300 python profiles/microbench/degenerate_lz4.py > gpurun_out/lz4_degenerate_r2v.txt 2>&1
timeout 300 python profiles/microbench/blocklen_lz4.py > gpurun_out/lz4_blocklen_r2v.txt 2>&1
for c in 4 3; do for s in 0 1; do timeout 300 python profiles/tools/jfif_time.py --dim 16384 --sub $s --comp $c --iters 4 --e2e 2>&1 | tail -2; done; done > gpurun_out/jfif_timing_r2v.txt
for s in 0 1; do timeout 300 python profiles/tools/jfif_time.py --dim 16384 --sub $s --iters 4 --natural 2>&1 | tail -1; done >> gpurun_out/jfif_timing_r2v.txt
tail -3 gpurun_out/lz4_phases_r2v.txt; cat gpurun_out/lz4_degenerate_r2v.txt gpurun_out/lz4_blocklen_r2v.txt gpurun_out/jfif_timing_r2v.txt
tail -c 300 gpurun_out/bench_r2v.err; head -c 600 gpurun_out/bench_r2v.json; echo; head -c 900 gpurun_out/bench_r2v_reference.json
