set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 300 python -m pytest tests/test_gpu_lz4.py -x -q 2>&1 | tail -1
timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tail -1
timeout 300 python profiles/microbench/degenerate_lz4.py 2>&1 | tail -6
cd lz4-jpeg_b200
for v in "4 8 4" "4 8 12" "4 8 16" "4 12 8" "4 16 8"; do
  set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DLJB_TINYLIST=$1 -DLJB_TINYMAX=$2 -DLJB_TINYCROWD=$3 -c csrc/lz4_encode.cu -o build/lz4_encode.cu.o || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o liblz4jpeg_b200.so build/*.o -lcudart || exit 1
  echo "variant tiny=$1 max=$2 crowd=$3"
  (cd .. && timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tail -1; timeout 300 python profiles/microbench/degenerate_lz4.py 2>&1 | grep "random\|text")
done
