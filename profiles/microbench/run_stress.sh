# repeats the parity suites of the kernels that changed most this round (no compute-sanitizer on this pool): a race would show up as a flaky byte
set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
fail=0
for i in $(seq 1 12); do
  timeout 300 python -m pytest tests/test_gpu_jfif.py tests/test_gpu_lz4.py -x -q -p no:cacheprovider 2>&1 | tail -1 | grep -q "passed" || fail=$((fail+1))
done
echo "stress: 12 repetitions of test_gpu_jfif.py + test_gpu_lz4.py, $fail failed"
for i in $(seq 1 6); do timeout 120 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench parity', d['parity_sample']['mismatch'], d['jpeg']['parity_sample']['mismatch'], d['jfif']['parity_sample']['mismatch'], d['jfif']['stb_rule_420']['parity_sample']['mismatch'], d['batch']['parity_sample']['mismatch'], d['jpeg']['e2e']['stream_equals_device_resident_stream'], d['jfif']['e2e']['file_equals_device_resident_file'], d['jfif']['stb_rule_420']['e2e']['file_equals_device_resident_file'])
"; done
