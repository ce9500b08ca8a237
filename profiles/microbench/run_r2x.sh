set -x
python -c "import lz4jpeg_b200 as l; l._native.lib()" || exit 1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
