set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 600 python -m pytest tests/test_gpu_lz4.py tests/test_gpu_batch.py -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-sample > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; echo bench rc=$?; tail -c 600 gpurun_out/bench_r2j.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2j.json').read().strip().splitlines()[-1])
print("lz4", d["value"], "e2e", d["e2e"]["value"], "decode", d["lz4_decode"]["value"], d["lz4_decode"]["roundtrip"])
print("batch", d["batch"]["value"], d["batch"]["kernels_ms"], "e2e", d["batch"]["e2e"]["value"])
PY
