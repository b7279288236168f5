set -x
mkdir -p gpurun_out
timeout 300 python scripts/e2e_probe.py 2>&1 | tail -4
timeout 600 python -m pytest tests -m gpu -x -q -k "shadows or any or trace_rays" 2>&1 | tail -4
