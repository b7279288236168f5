set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dynamic.py tests/test_gpu_parity.py -m gpu -x -q -k "any or shadows or skin" 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python scripts/time_next.py 2>&1 | tail -20
