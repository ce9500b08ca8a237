set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
for v in "" "-DLJB_JFIF_I2F"; do
(cd lz4-jpeg_b200 && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $v -c csrc/jfif_encode.cu -o build/jfif_encode.cu.o && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o liblz4jpeg_b200.so build/*.o -lcudart) || exit 1
echo "variant $v"
timeout 600 python -m pytest tests/test_gpu_jfif.py -x -q 2>&1 | tail -1
timeout 300 python profiles/tools/jfif_time.py --dim 16384 --sub 0 --iters 5 2>&1 | tail -1
timeout 300 python profiles/tools/jfif_time.py --dim 16384 --sub 1 --iters 5 2>&1 | tail -1
done
