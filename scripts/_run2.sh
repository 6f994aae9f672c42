timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
python scripts/perf_c1.py | tail -1
