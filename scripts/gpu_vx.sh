python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python scripts/run_configs.py config5 2>&1 | cut -c1-900
