(timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -3) > gpurun_out/c17_bench.log 2>&1
cat gpurun_out/c17_bench.log | cut -c1-3000
(SKB_ADJ_MODE=1 timeout 300 python tools/time_bwd.py cfg4 2>&1 | tail -4)
(timeout 900 python -m pytest tests/test_gpu_backward.py -q 2>&1 | tail -3)
