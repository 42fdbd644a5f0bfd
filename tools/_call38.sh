timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:fwd5_kernel<.*\(int\)6, \(int\)32>' -s 1 -c 1 -o gpurun_out/r02_bwd_sym_cfg4 -f python tools/run_sym.py 3 > gpurun_out/c38_ncu.log 2>&1
tail -2 gpurun_out/c38_ncu.log
