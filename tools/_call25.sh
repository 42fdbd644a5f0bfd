timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:fwd5_kernel<.*\(int\)[45], \(int\)(16|32)>' -s 2 -c 2 -o gpurun_out/r02_bwd_cfg4c -f python tools/run_cfg.py cfg4f 2 bwd > gpurun_out/c25_ncu.log 2>&1
tail -3 gpurun_out/c25_ncu.log
