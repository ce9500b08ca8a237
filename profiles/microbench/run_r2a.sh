# round 2, first GPU run: the chain-search LZ4 kernel against the full search; tests; phases; degenerate inputs; decoder timing
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tee gpurun_out/lz4_lazy_r2a.txt
LJB_LZ4_SEARCH=full timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tee gpurun_out/lz4_full_r2a.txt
LJB_LZ4_PHASES=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_phases_r2a.txt 2>&1; tail -4 gpurun_out/lz4_phases_r2a.txt
timeout 300 python profiles/microbench/degenerate_lz4.py 2>&1 | tee gpurun_out/lz4_degenerate_r2a.txt
timeout 300 python profiles/microbench/blocklen_lz4.py 2>&1 | tee gpurun_out/lz4_blocklen_r2a.txt
timeout 200 python profiles/microbench/quick_lz4_decode.py 268435456 2>&1 | tee gpurun_out/lz4_decode_r2a.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_encode -s 1 -c 1 -f -o gpurun_out/lz4_r2a python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_ncu.log 2>&1; tail -2 gpurun_out/lz4_ncu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo bench rc=$?; tail -c 600 gpurun_out/bench_r2a.err; head -c 1500 gpurun_out/bench_r2a.json
