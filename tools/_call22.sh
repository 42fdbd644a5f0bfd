(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 2>&1 | tail -2) > gpurun_out/r02_bench_n8.json 2>&1
cut -c1-700 gpurun_out/r02_bench_n8.json
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('cfg5_sharded'))
PY
