timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "phased or pipelined" 2>&1 | tail -15
