timeout 900 python -m pytest tests/test_gpu_backward.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/time_bwd.py cfg4 2>&1 | tail -3
