python -m pytest tests/test_gpu_epilogue.py -m gpu -q -x -k "unreached or refuses" 2>&1 | tail -40
