timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_case.py > gpurun_out/c40_memcheck.log 2>&1; tail -4 gpurun_out/c40_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_case.py > gpurun_out/c40_racecheck.log 2>&1; tail -4 gpurun_out/c40_racecheck.log
