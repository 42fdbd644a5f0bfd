(timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -E "passed|failed|FAILED|Error|assert [0-9n]" | head -20) > gpurun_out/c32_tests.log 2>&1
cat gpurun_out/c32_tests.log
(timeout 300 python tools/time_bwd.py cfg4 cfg3b 2>&1 | tail -8) > gpurun_out/c32_time_bwd.log 2>&1
cat gpurun_out/c32_time_bwd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd5_kernel -s 2 -c 1 -o gpurun_out/r02_fwd5_cfg3 -f python tools/run_cfg.py cfg3 4 > gpurun_out/c32_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:fwd5_kernel<.*\(int\)[45], \(int\)(16|32)>' -s 2 -c 2 -o gpurun_out/r02_bwd_cfg4 -f python tools/run_cfg.py cfg4f 2 bwd > gpurun_out/c32_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deriv_stream_kernel -s 1 -c 1 -o gpurun_out/r02_deriv_cfg3 -f python tools/time_deriv.py > gpurun_out/c32_ncu3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/c32_benchncu.log 2>&1
(timeout 900 python bench.py 2>&1 | tail -1) > gpurun_out/r02_bench_n1.json
cut -c1-600 gpurun_out/r02_bench_n1.json
(timeout 300 python tools/time_deriv.py 2>&1 | tail -2) > gpurun_out/c32_deriv.log; cat gpurun_out/c32_deriv.log
ls -la gpurun_out | tail -8
