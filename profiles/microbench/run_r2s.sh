set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 900 python -m pytest tests/test_gpu_jfif.py -x -q 2>&1 | tail -5
timeout 300 python profiles/tools/jfif_time.py --dim 16384 --sub 0 --iters 5 2>&1 | tail -1
timeout 300 python profiles/tools/jfif_time.py --dim 16384 --sub 1 --iters 5 2>&1 | tail -1
timeout 300 python profiles/tools/jfif_time.py --dim 16384 --sub 0 --iters 5 --natural 2>&1 | tail -1
timeout 300 python profiles/tools/jfif_time.py --dim 16384 --sub 1 --iters 5 --natural 2>&1 | tail -1
