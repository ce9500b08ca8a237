set -x
nvidia-smi --query-gpu=index,name --format=csv
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 600 python -m pytest tests/test_gpu_compat.py tests/test_gpu_sharding.py -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r2f_2gpu.json 2> gpurun_out/bench_r2f_2gpu.err; echo bench rc=$?; tail -c 800 gpurun_out/bench_r2f_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2f_2gpu.json').read().strip().splitlines()[-1])
print("lz4", d["value"], "e2e", d["e2e"]["value"], "ceiling", d["e2e"]["host_ceiling_gbs"], "strong", d["strong_scaling"])
print("jpeg", d["jpeg"]["value"], "e2e", d["jpeg"]["e2e"]["value"], "ceiling", d["jpeg"]["e2e"]["host_ceiling_mpix"])
print("batch", d["batch"]["value"], "e2e", d["batch"]["e2e"]["value"])
print("parity", d["parity_sample"], d["jpeg"]["parity_sample"])
PY
