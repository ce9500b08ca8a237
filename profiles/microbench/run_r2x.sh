set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
cd lz4-jpeg_b200
for v in "-DLJB_EXTEND_MERGED=1" "-DLJB_EXTEND_MERGED=0" "-DLJB_EXTEND_MERGED=1" "-DLJB_EXTEND_MERGED=0"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $v -c csrc/lz4_encode.cu -o build/lz4_encode.cu.o 2>/dev/null || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o liblz4jpeg_b200.so build/*.o -lcudart || exit 1
  echo "variant $v"
  (cd .. && timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tail -2)
done
