(timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -2)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/final_benchncu.log 2>&1
(timeout 900 python bench.py 2>&1 | tail -1) > gpurun_out/r02_bench_n1.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read())
print("N=1", d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['cfg4']['ms_per_step'], d['cfg4']['gram_with_grad_points_ms'], d['clocks'])
PY
