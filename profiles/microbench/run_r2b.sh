# round 2, second GPU run: chain search with interleaved / contiguous walker segments, three-tier evaluation
set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tee gpurun_out/lz4_lazy_r2b.txt
LJB_LZ4_TUNE=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tee gpurun_out/lz4_lazy_contig_r2b.txt
LJB_LZ4_PHASES=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_phases_r2b.txt 2>&1; tail -4 gpurun_out/lz4_phases_r2b.txt
LJB_LZ4_TUNE=1 LJB_LZ4_PHASES=1 timeout 120 python profiles/microbench/quick_lz4.py 268435456 2>&1 | tail -3
timeout 300 python profiles/microbench/degenerate_lz4.py 2>&1 | tee gpurun_out/lz4_degenerate_r2b.txt
timeout 200 python profiles/microbench/quick_lz4_decode.py 268435456 2>&1 | tee gpurun_out/lz4_decode_r2b.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_encode -s 1 -c 1 -f -o gpurun_out/lz4_r2b python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_ncu.log 2>&1; tail -2 gpurun_out/lz4_ncu.log
