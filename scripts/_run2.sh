timeout 600 python -m pytest tests/test_gpu_integrate.py -m gpu -q -x -k "time_dependent" 2>&1 | tail -30
