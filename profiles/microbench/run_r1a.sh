set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; echo bench rc=$?
python profiles/microbench/quick_jpeg.py 4096 > gpurun_out/jpeg_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:jpeg_encode -s 1 -c 1 -f -o gpurun_out/jpeg_v1 python profiles/microbench/quick_jpeg.py 4096 > gpurun_out/jpeg_ncu.log 2>&1
cat gpurun_out/jpeg_plain.log
