set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
cd lz4-jpeg_b200
for v in 0 4 8; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DLJB_DEC_AHEAD=$v -c csrc/lz4_decode.cu -o build/lz4_decode.cu.o 2>/dev/null || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o liblz4jpeg_b200.so build/*.o -lcudart || exit 1
  echo "variant ahead=$v"
  (cd .. && LJB_ALLOW_PHANTOM=1 timeout 300 python profiles/microbench/quick_lz4_decode.py 2147483648 2>&1 | grep "decode (device)" | tail -1)
done
