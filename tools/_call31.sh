timeout 600 python -m pytest tests/test_gpu_derivatives.py -x -q -m gpu 2>&1 | tail -12
timeout 300 python tools/time_deriv.py 2>&1 | tail -3
