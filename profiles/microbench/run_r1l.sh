# round-1 final evidence run (r1l): tests, smoke, bench + reference arm, launch list with DRAM traffic, ncu --set full of the LZ4 kernel
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err; echo bench rc=$?
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r1l_reference.json 2>> gpurun_out/bench_r1l.err; echo ref rc=$?
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_launch.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1l.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lz4_encode -s 1 -c 1 -f -o gpurun_out/lz4_r1l python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_ncu.log 2>&1
LJB_LZ4_PHASES=1 python profiles/microbench/quick_lz4.py 268435456 > gpurun_out/lz4_phases_r1l.txt 2>&1
python profiles/microbench/degenerate_lz4.py > gpurun_out/lz4_degenerate_r1l.txt 2>&1
cat gpurun_out/lz4_plain.log gpurun_out/lz4_degenerate_r1l.txt
tail -c 400 gpurun_out/bench_r1l.err
