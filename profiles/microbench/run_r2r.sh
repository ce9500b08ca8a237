# source-level captures of the two JPEG encode kernels (JFIF 4:4:4 and the JPEG-like encoder)
set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 300 python profiles/tools/jfif_time.py --dim 8192 --sub 0 --iters 3 2>&1 | tail -3
timeout 300 python profiles/microbench/quick_jpeg.py 8192 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jfif_encode -s 1 -c 1 -f -o gpurun_out/jfif444_r2r python profiles/tools/jfif_time.py --dim 8192 --sub 0 --iters 2 > gpurun_out/jfif_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jpeg_encode -s 1 -c 1 -f -o gpurun_out/jpeg_r2r python profiles/microbench/quick_jpeg.py 8192 > gpurun_out/jpeg_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
