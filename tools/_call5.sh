set -x
(timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_tile.py -x -q 2>&1 | tail -5) > gpurun_out/c6_tests.log 2>&1
cat gpurun_out/c6_tests.log
(timeout 300 python tools/time_tile.py cfg3 cfg5s cfg4f 2>&1 | tail -30) > gpurun_out/c6_time_tile.log 2>&1
cat gpurun_out/c6_time_tile.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd5_kernel -s 1 -c 1 -o gpurun_out/c6_fwd5_cfg3 -f env SKB_TILE_MODE=0 python tools/run_cfg.py cfg3 3 > gpurun_out/c6_ncu.log 2>&1
tail -2 gpurun_out/c6_ncu.log
