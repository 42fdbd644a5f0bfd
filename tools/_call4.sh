set -x
(timeout 300 python tools/time_wpsm.py 2>&1 | tail -12) > gpurun_out/c4_wpsm.log 2>&1
cat gpurun_out/c4_wpsm.log
