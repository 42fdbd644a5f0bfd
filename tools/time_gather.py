"""Multi-GPU gather micro-benchmark (development aid, run under torchrun): symmetric-memory barrier vs NCCL barrier vs
all_gather_into_tensor vs the solver with peer stores, per step, max over ranks."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb  # noqa: E402


def timed(fn, n=50):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import torch.distributed._symmetric_memory as symm_mem
    A = 128
    g = torch.Generator().manual_seed(rank)
    X = torch.rand((world * A, 64, 5), dtype=torch.float64, generator=g).cuda()
    Y = torch.rand((A, 64, 5), dtype=torch.float64, generator=g).cuda()
    sk = skb.SigKernel(skb.RBFKernel(0.5), 2)
    t = symm_mem.empty((world * A, A), dtype=torch.float64, device="cuda")
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    blk = torch.empty((A, A), dtype=torch.float64, device="cuda")
    full = torch.empty((world * A, A), dtype=torch.float64, device="cuda")
    lo = rank * A
    ptrs = [int(q) + lo * A * 8 for q in hdl.buffer_ptrs]
    res = {
        "symm barrier": timed(lambda: hdl.barrier(channel=0)),
        "nccl barrier": timed(lambda: dist.barrier()),
        "all_gather 128KB": timed(lambda: dist.all_gather_into_tensor(full, blk)),
        "solve local": timed(lambda: skb.ops.sigkernel_forward(X[lo:lo + A], Y, "rbf", 0.5, 2, "gram")),
        "solve peers (no barrier)": timed(lambda: skb.ops.sigkernel_forward_peers(X[lo:lo + A], Y, "rbf", 0.5, 2, ptrs, "gram")),
        "solve peers + symm barrier": timed(lambda: (skb.ops.sigkernel_forward_peers(X[lo:lo + A], Y, "rbf", 0.5, 2, ptrs, "gram"), hdl.barrier(channel=0))),
        "solve local + all_gather": timed(lambda: dist.all_gather_into_tensor(full, skb.ops.sigkernel_forward(X[lo:lo + A], Y, "rbf", 0.5, 2, "gram"))),
        "compute_Gram_sharded": timed(lambda: skb.distributed.compute_Gram_sharded(sk, X, Y)),
    }
    if rank == 0:
        for k, v in res.items():
            print(f"N={world} {k}: {v*1e3:.1f} us", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
