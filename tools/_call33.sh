for mode in kernel handle kernel handle; do
(SKB_RANK_BARRIER=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 50 --warmup 5 2>&1 | tail -1) > gpurun_out/c33_$mode.json 2>&1
python - <<PY
import json
d=json.loads(open('gpurun_out/c33_$mode.json').read())
print("$mode", d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['cfg5_sharded']['ms_per_step'])
PY
done
