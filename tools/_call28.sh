timeout 200 python tools/probe_fp64_latency.py 2>&1 | tail -16
SKB_WPSM=0 timeout 300 python tools/time_fwd.py cfg3 2>&1 | grep rbf
timeout 300 python tools/time_bwd.py cfg4 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q -m gpu 2>&1 | tail -2
