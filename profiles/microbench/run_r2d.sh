set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo bench rc=$?; tail -c 1500 gpurun_out/bench_r2d.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2d.json').read().strip().splitlines()[-1])
def show(o,ind=0):
    for k,v in o.items():
        if isinstance(v,dict): print(' '*ind+k+':'); show(v,ind+2)
        else: print(' '*ind+f"{k}: {str(v)[:160]}")
show(d)
PY
