(timeout 1500 python -m pytest tests/test_gpu_backward.py -x -q 2>&1 | tail -25) > gpurun_out/c10_tests.log 2>&1
cat gpurun_out/c10_tests.log
python tools/recon_probe.py cfg4 cfg3b 2>&1 | tail -3
(timeout 300 python tools/time_bwd.py cfg4 cfg3b 2>&1 | tail -12) > gpurun_out/c10_time_bwd.log 2>&1
cat gpurun_out/c10_time_bwd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd5_kernel -s 1 -c 1 -o gpurun_out/c10_recon_l16 -f python tools/recon_probe.py cfg4 > gpurun_out/c10_ncu.log 2>&1
SKB_ADJ_MODE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd5_kernel -s 1 -c 1 -o gpurun_out/c10_recon_l32 -f python tools/recon_probe.py cfg4 >> gpurun_out/c10_ncu.log 2>&1
tail -3 gpurun_out/c10_ncu.log
