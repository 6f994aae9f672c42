timeout 600 python -m pytest tests/test_gpu_joint.py -m gpu -q -x 2>&1 | tail -15
