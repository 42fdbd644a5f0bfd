(timeout 600 python -m pytest tests/test_gpu_multi.py -q -s 2>&1 | tail -8) > gpurun_out/c18_multi.log 2>&1
cat gpurun_out/c18_multi.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | tail -2) > gpurun_out/c18_bench2.log 2>&1
cut -c1-2500 gpurun_out/c18_bench2.log
