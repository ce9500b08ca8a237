# round-1 re-entry evidence run: tests, smoke, bench + reference arm, launch list, ncu --set full of the JFIF kernel
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; echo bench rc=$?
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r1e_reference.json 2>> gpurun_out/bench_r1e.err; echo ref rc=$?
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_launch.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
for s in 0 1; do python profiles/tools/jfif_time.py --sub $s; python profiles/tools/jfif_time.py --sub $s --natural; done > gpurun_out/jfif_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jfif_encode -s 1 -c 1 -f -o gpurun_out/jfif_r1e python profiles/tools/jfif_time.py --sub 1 --iters 1 > gpurun_out/jfif_ncu.log 2>&1
cat gpurun_out/jfif_plain.log
tail -c 600 gpurun_out/bench_r1e.err
