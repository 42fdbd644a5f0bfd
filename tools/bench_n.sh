# usage: bash tools/bench_n.sh N   (under gpurun --gpus N): the bench line at N GPUs -> gpurun_out/r02_bench_nN.json
N=${1:-2}
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 5 2>&1 | tail -1) > gpurun_out/r02_bench_n$N.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n$N.json').read())
print("N=$N", d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['cfg5_sharded']['ms_per_step'])
PY
