set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_fwd -s 1 -c 1 -o gpurun_out/c2_tile_cfg3 -f python tools/run_cfg.py cfg3 3 > gpurun_out/c2_ncu.log 2>&1
tail -3 gpurun_out/c2_ncu.log
ls -la gpurun_out/*.ncu-rep
