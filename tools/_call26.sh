SKB_WPSM=0 timeout 300 python tools/time_fwd.py cfg3 cfg4f cfg5s 2>&1 | grep rbf
timeout 300 python tools/time_bwd.py cfg4 2>&1 | tail -3
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
