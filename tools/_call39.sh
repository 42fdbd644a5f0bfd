timeout 900 python -m pytest tests/test_gpu_backward.py -x -q -m gpu 2>&1 | tail -2
bash tools/_call37.sh
