set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_post.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python scripts/time_post.py 2>&1 | tail -5
