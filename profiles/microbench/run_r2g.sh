set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python profiles/microbench/blocklen_lz4.py 2>&1 | tee gpurun_out/lz4_blocklen_r2g.txt
LJB_LZ4_NO_SMALL=1 timeout 300 python profiles/microbench/blocklen_lz4.py 2>&1 | tee gpurun_out/lz4_blocklen_nosmall_r2g.txt
