# end of round 2: the bench at 8 GPUs (weak + strong scaling, end to end with three-byte pixels)
set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
for n in 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2ai_${n}gpu.json 2> gpurun_out/bench_r2ai_${n}gpu.err; echo bench $n rc=$?; tail -c 400 gpurun_out/bench_r2ai_${n}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2ai_${n}gpu.json').read().strip().splitlines()[-1])
print("N=$n lz4", d["value"], "e2e", d["e2e"]["value"], "ceiling", d["e2e"]["host_ceiling_gbs"], "decode", d["lz4_decode"]["value"])
print("  strong", d["strong_scaling"])
print("  jpeg", d["jpeg"]["value"], "e2e", d["jpeg"]["e2e"]["value"], "ceiling", d["jpeg"]["e2e"]["host_ceiling_mpix"], "rgba", d["jpeg"]["e2e_rgba"]["value"], d["jpeg"]["e2e_rgba"]["host_ceiling_mpix"])
print("  jfif", d["jfif"]["value"], "e2e", d["jfif"]["e2e"]["value"], d["jfif"]["e2e_rgba"]["value"], "420", d["jfif"]["stb_rule_420"]["value"], d["jfif"]["stb_rule_420"]["e2e"]["value"])
print("  batch", d["batch"]["value"], "e2e", d["batch"]["e2e"]["value"])
print("  parity", d["parity_sample"], d["jpeg"]["parity_sample"], d["batch"]["parity_sample"])
PY
done
