timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -s 2>&1 | tail -8
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 30 --warmup 3 2>&1 | tail -1) > gpurun_out/c30_n2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c30_n2.json').read())
print("N=2", d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['cfg5_sharded']['ms_per_step'])
PY
