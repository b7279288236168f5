python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for p in 0 4 8 16 32; do echo "== PREFETCH $p"; IDKPT_TUNE_PREFETCH=$p python scripts/per_bounce.py; done
