(timeout 300 python tools/time_tile.py cfg3 cfg5s 2>&1 | tail -30) > gpurun_out/c7_time_tile.log 2>&1
cat gpurun_out/c7_time_tile.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_fwd -s 1 -c 1 -o gpurun_out/c7_tile_cfg3 -f python tools/run_cfg.py cfg3 3 > gpurun_out/c7_ncu.log 2>&1
tail -2 gpurun_out/c7_ncu.log
