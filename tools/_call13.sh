(timeout 1500 python -m pytest tests/test_gpu_backward.py -q 2>&1 | grep -E "passed|failed|FAILED|assert [0-9]" | head -40) > gpurun_out/c13_tests.log 2>&1
cat gpurun_out/c13_tests.log
(timeout 300 python tools/time_bwd.py cfg4 cfg3b 2>&1 | tail -12) > gpurun_out/c13_time_bwd.log 2>&1
cat gpurun_out/c13_time_bwd.log
