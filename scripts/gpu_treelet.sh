python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for p in 0 192 384 512 768; do echo "== TREELET_PAIRS $p"; IDKPT_TREELET_PAIRS=$p python scripts/per_bounce.py; done
