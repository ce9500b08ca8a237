set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 600 python -m pytest tests/test_gpu_lz4.py tests/test_gpu_sharding.py -x -q 2>&1 | tail -3
for t in 0 2 3; do echo "== tune $t (walker segment bytes: 0 -> 33, 2 -> 66, 3 -> 44)"; LJB_LZ4_TUNE=$t timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tail -1; LJB_LZ4_TUNE=$t LJB_LZ4_PHASES=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | grep "ljb lz4" | tail -2; done
timeout 300 python profiles/microbench/degenerate_lz4.py 2>&1 | tee gpurun_out/lz4_degenerate_r2h.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_encode -s 1 -c 1 -f -o gpurun_out/lz4_r2h python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_ncu.log 2>&1; tail -2 gpurun_out/lz4_ncu.log
