# source-level capture of the JFIF 4:4:4 encode kernel
set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 300 python profiles/tools/jfif_time.py --dim 8192 --sub 0 --iters 3 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jfif_encode -s 1 -c 1 -f -o gpurun_out/jfif444_r2s python profiles/tools/jfif_time.py --dim 8192 --sub 0 --iters 2 > gpurun_out/jfif_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jfif_encode -s 1 -c 1 -f -o gpurun_out/jfif444nat_r2s python profiles/tools/jfif_time.py --dim 8192 --sub 0 --iters 2 --natural > gpurun_out/jfif_ncu2.log 2>&1
