# rebuilds lz4_encode.cu on the GPU box with different tier thresholds of the chain search and times 256 MiB of text
set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
cd lz4-jpeg_b200
for v in "0 32 32" "1 32 32" "2 32 32" "3 32 32" "2 24 32" "2 16 32"; do
  set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DLJB_TINYLIST=$1 -DLJB_BIGLIST=$2 -DLJB_VLONG=$3 -c csrc/lz4_encode.cu -o build/lz4_encode.cu.o || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o liblz4jpeg_b200.so build/*.o -lcudart || exit 1
  echo "variant tiny=$1 big=$2 vlong=$3"
  (cd .. && timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tail -1)
done
