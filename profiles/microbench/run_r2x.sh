set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 600 python -m pytest tests/test_gpu_lz4.py tests/test_gpu_inverse_chain.py tests/test_gpu_batch.py -x -q 2>&1 | tail -1
cd lz4-jpeg_b200
for v in "-DLJB_FUSE_ROUNDS=1" "-DLJB_FUSE_ROUNDS=0" "-DLJB_FUSE_ROUNDS=1" "-DLJB_FUSE_ROUNDS=0"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $v -c csrc/lz4_encode.cu -o build/lz4_encode.cu.o 2>/dev/null || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o liblz4jpeg_b200.so build/*.o -lcudart || exit 1
  echo "variant $v"
  (cd .. && timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tail -2; LJB_LZ4_PHASES=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | grep "cycles per block" | cut -c1-220)
done
