python scripts/perf_r2.py 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
