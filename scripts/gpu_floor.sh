python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for l2 in 1 0; do for v in 3 2; do echo "== L2_PERSIST $l2 VARIANT $v"; IDKPT_L2_PERSIST=$l2 IDKPT_TRAVERSE_VARIANT=$v python scripts/per_bounce.py; done; done
