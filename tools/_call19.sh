(timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -E "passed|failed|FAILED|Error|assert [0-9n]" | head -20) > gpurun_out/c19_tests.log 2>&1
cat gpurun_out/c19_tests.log
(timeout 300 python tools/time_bwd.py cfg4 cfg3b 2>&1 | tail -8) > gpurun_out/c19_time_bwd.log 2>&1
cat gpurun_out/c19_time_bwd.log
(timeout 900 python baseline/time_ref_numba.py gpurun_out/r02_ref_numba_b200.json 2>&1 | tail -3 | cut -c1-1500)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
