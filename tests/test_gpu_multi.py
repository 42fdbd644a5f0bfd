"""2-GPU NCCL test of the sharded Gram (runs only where >= 2 CUDA devices are visible)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sigkernel_b200 as skb
        g = torch.Generator().manual_seed(0)
        X = torch.rand((13, 20, 3), dtype=torch.float64, generator=g).cuda()
        Y = torch.rand((9, 17, 3), dtype=torch.float64, generator=g).cuda()
        sk = skb.SigKernel(skb.RBFKernel(0.5), 1)
        G = skb.distributed.compute_Gram_sharded(sk, X, Y)
        ref = sk.compute_Gram(X, Y)
        q.put((rank, bool(torch.equal(G, ref))))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_gram_two_gpus_nccl():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert dict(q.get(timeout=5) for _ in range(2)) == {0: True, 1: True}
