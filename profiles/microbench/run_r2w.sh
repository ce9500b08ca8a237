set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_encode -s 1 -c 1 -f -o gpurun_out/lz4_r2w python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_ncu.log 2>&1
