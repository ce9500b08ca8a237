set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
for c in 4 3; do for s in 0 1; do timeout 300 python profiles/tools/jfif_time.py --dim 16384 --sub $s --comp $c --iters 4 --e2e 2>&1 | tail -2; done; done
