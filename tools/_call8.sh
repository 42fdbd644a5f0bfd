(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25) > gpurun_out/c8_tests.log 2>&1
cat gpurun_out/c8_tests.log
(timeout 300 python tools/time_bwd.py cfg4 cfg3b 2>&1 | tail -12) > gpurun_out/c8_time_bwd.log 2>&1
cat gpurun_out/c8_time_bwd.log
