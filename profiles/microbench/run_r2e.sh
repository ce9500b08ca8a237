# round 2: tests incl. batch / process / comm / sharding / >4 GiB, compute-sanitizer logs, bench with the batch object
set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
mkdir -p gpurun_out/sanitizer
for tool in memcheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python profiles/microbench/sanitize_small.py > gpurun_out/sanitizer/${tool}_r2e.log 2>&1; echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer/${tool}_r2e.log
done
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python profiles/microbench/sanitize_small.py small > gpurun_out/sanitizer/racecheck_r2e.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/sanitizer/racecheck_r2e.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo bench rc=$?; tail -c 1500 gpurun_out/bench_r2e.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2e.json').read().strip().splitlines()[-1])
def show(o,ind=0):
    for k,v in o.items():
        if isinstance(v,dict): print(' '*ind+k+':'); show(v,ind+2)
        else: print(' '*ind+f"{k}: {str(v)[:200]}")
show(d.get('batch') or {})
PY
