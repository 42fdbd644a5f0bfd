set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_tile.py -x -q 2>&1 | tail -15) > gpurun_out/c3_tile_tests.log 2>&1
cat gpurun_out/c3_tile_tests.log
(timeout 300 python tools/time_tile.py cfg3 cfg5s 2>&1 | tail -12) > gpurun_out/c3_time_tile.log 2>&1
cat gpurun_out/c3_time_tile.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_fwd -s 1 -c 1 -o gpurun_out/c3_tile_cfg3 -f python tools/run_cfg.py cfg3 3 > gpurun_out/c3_ncu.log 2>&1
tail -2 gpurun_out/c3_ncu.log
