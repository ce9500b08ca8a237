set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 600 python -m pytest tests/test_gpu_lz4.py -x -q 2>&1 | tail -3
for t in 0 64 256 4096 2048 4160 2112; do echo "== tune $t"; LJB_LZ4_TUNE=$t timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tail -1; done
LJB_LZ4_PHASES=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_phases_r2c.txt 2>&1; tail -4 gpurun_out/lz4_phases_r2c.txt
timeout 300 python profiles/microbench/degenerate_lz4.py 2>&1 | tee gpurun_out/lz4_degenerate_r2c.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_encode -s 1 -c 1 -f -o gpurun_out/lz4_r2c python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_ncu.log 2>&1; tail -2 gpurun_out/lz4_ncu.log
