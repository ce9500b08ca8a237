set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
for ch in 134217728 67108864 50331648 33554432 16777216; do echo "chunk $ch"; LJB_PIPE_CHUNK_BYTES=$ch timeout 300 python profiles/microbench/quick_e2e_jpeg.py 2>&1 | tail -2; done
