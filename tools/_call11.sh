(timeout 1500 python -m pytest tests/test_gpu_backward.py -q 2>&1 | grep -E "passed|failed|FAILED|assert [0-9]" | head -60) > gpurun_out/c11_tests.log 2>&1
cat gpurun_out/c11_tests.log
