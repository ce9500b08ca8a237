set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
for t in 0 2 4; do echo "== tune $t (short lists up to: 0 -> 8, 2 -> 31, 4 -> 16)"; LJB_LZ4_TUNE=$t timeout 300 python -m pytest tests/test_gpu_lz4.py -x -q 2>&1 | tail -1; LJB_LZ4_TUNE=$t timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tail -1; LJB_LZ4_TUNE=$t LJB_LZ4_PHASES=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | grep "ljb lz4 phases" | tail -1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_encode -s 1 -c 1 -f -o gpurun_out/lz4_r2n python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_ncu.log 2>&1; tail -2 gpurun_out/lz4_ncu.log
