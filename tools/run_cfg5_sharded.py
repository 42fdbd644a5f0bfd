"""BASELINE.json configs[4]: compute_Gram 512x512 len=128 dim=8 dyadic_order=2 RBF sharded over the ranks of one
box (rows of X per rank, Y replicated, one NCCL all-gather of G).  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_cfg5_sharded.py

Prints one JSON line (rank 0): time = max over ranks of (solve + all-gather), CUDA events; the leading 3x3 block
is checked against the committed reference fixture tests/golden/cfg5_gram_rbf.npz."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(0)                     # same recipe as tests/golden/make_golden.py
    X = torch.rand((512, 128, 8), dtype=torch.float64, generator=g).to(dev)
    Y = torch.rand((512, 128, 8), dtype=torch.float64, generator=g).to(dev)
    sk = skb.SigKernel(skb.RBFKernel(0.5), 2)

    def run():
        if world > 1:
            return skb.distributed.compute_Gram_sharded(sk, X, Y)
        return sk.compute_Gram(X, Y)

    for _ in range(3):
        G = run()
    times = []
    for _ in range(5):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        G = run()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    if rank == 0:
        z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                 "cfg5_gram_rbf.npz"))
        ref = z["G"]
        got = G[:ref.shape[0], :ref.shape[1]].cpu().numpy()
        err = float(np.max(np.abs(got - ref) / (np.abs(ref) + 1.0)))
        best = min(times)
        print(json.dumps({"workload": "compute_Gram 512x512 len=128 dim=8 dyadic_order=2 RBF (BASELINE configs[4])",
                          "n_gpus": world, "ms_best": best, "ms_all": times, "pairs_per_s": 512 * 512 / (best * 1e-3),
                          "parity_leading_block_vs_reference": err}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
