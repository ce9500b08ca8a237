# round-2 evidence run: tests, smoke, bench + reference arm, launch list with DRAM traffic, ncu --set full of the LZ4 kernels
set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err; echo bench rc=$?
timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r2p_reference.json 2>> gpurun_out/bench_r2p.err; echo ref rc=$?
LJB_BENCH_BATCH_IMAGES=64 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-sample > gpurun_out/plain_launch.log 2>&1 &&
LJB_BENCH_BATCH_IMAGES=64 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2p.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-sample > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_encode -s 1 -c 1 -f -o gpurun_out/lz4_r2p python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_decode -s 1 -c 1 -f -o gpurun_out/lz4dec_r2p python profiles/microbench/quick_lz4_decode.py 268435456 > gpurun_out/lz4dec_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_small -s 1 -c 1 -f -o gpurun_out/lz4small_r2p python profiles/microbench/blocklen_lz4.py > gpurun_out/lz4small_ncu.log 2>&1
LJB_LZ4_PHASES=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_phases_r2p.txt 2>&1
timeout 300 python profiles/microbench/degenerate_lz4.py > gpurun_out/lz4_degenerate_r2p.txt 2>&1
timeout 300 python profiles/microbench/blocklen_lz4.py > gpurun_out/lz4_blocklen_r2p.txt 2>&1
tail -3 gpurun_out/lz4_phases_r2p.txt; cat gpurun_out/lz4_degenerate_r2p.txt gpurun_out/lz4_blocklen_r2p.txt
tail -c 300 gpurun_out/bench_r2p.err; head -c 600 gpurun_out/bench_r2p.json; echo; head -c 900 gpurun_out/bench_r2p_reference.json
