set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python profiles/microbench/quick_lz4_decode.py 268435456 2>&1 | tee gpurun_out/lz4_decode_r2i.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; echo bench rc=$?; tail -c 600 gpurun_out/bench_r2i.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2i.json').read().strip().splitlines()[-1])
print("lz4", d["value"], "e2e", d["e2e"]["value"], "decode", d["lz4_decode"]["value"], d["lz4_decode"]["roundtrip"])
print("batch", d["batch"]["value"], d["batch"]["kernels_ms"], "e2e", d["batch"]["e2e"]["value"])
PY
