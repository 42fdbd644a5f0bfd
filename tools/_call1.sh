set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(timeout 600 python -m pytest tests/test_gpu_tile.py -x -q 2>&1 | tail -15) > gpurun_out/c1_tile_tests.log 2>&1
cat gpurun_out/c1_tile_tests.log
(timeout 300 python tools/time_tile.py cfg3 cfg5s 2>&1 | tail -12) > gpurun_out/c1_time_tile.log 2>&1
cat gpurun_out/c1_time_tile.log
(timeout 600 python baseline/time_ref_numba.py gpurun_out/ref_numba_b200.json 2>&1 | tail -5) > gpurun_out/c1_numba.log 2>&1
tail -c 1500 gpurun_out/c1_numba.log
