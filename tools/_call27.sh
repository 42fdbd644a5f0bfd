SKB_WPSM=0 timeout 300 python tools/time_fwd.py cfg3 2>&1 | grep rbf
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
(timeout 900 python bench.py --steps 30 --warmup 3 2>&1 | tail -1) > gpurun_out/c27_n1.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c27_n1.json').read())
print("N=1", d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['cfg4']['ms_per_step'], d['cfg4']['gram_with_grad_points_ms'])
PY
