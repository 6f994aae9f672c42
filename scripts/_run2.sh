python scripts/perf_r2.py k3 k3long 2>&1 | tail -1
GALAX_B200_LIB=build_variants/libgx_coef.so python scripts/perf_r2.py k3 k3long 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
