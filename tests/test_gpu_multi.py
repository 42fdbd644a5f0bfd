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
        ok = []
        # ragged split (13 rows over 2 ranks): padded NCCL all-gather
        G = skb.distributed.compute_Gram_sharded(sk, X, Y)
        ok.append(bool(torch.equal(G, sk.compute_Gram(X, Y))) and skb.distributed.last_gather == "all_gather")
        # even split: the solver stores straight into every rank's symmetric-memory copy of G, no collective; several calls
        # in a row alternate between the two copies
        for k in range(5):
            Xk = X[:12] * (1.0 + 0.01 * k)
            G = skb.distributed.compute_Gram_sharded(sk, Xk, Y)
            ref = sk.compute_Gram(Xk, Y)
            ok.append(bool(torch.equal(G, ref)))
        path = skb.distributed.last_gather
        # with gradients the NCCL path (and its autograd-aware gather) is taken
        Xg = X[:12].clone().requires_grad_(True)
        Gg = skb.distributed.compute_Gram_sharded(sk, Xg, Y)
        Gg.sum().backward()
        lo, hi = skb.distributed.row_block(12, rank, world)
        Xr = X[:12].clone().requires_grad_(True)
        sk.compute_Gram(Xr, Y).sum().backward()
        ok.append(bool(torch.allclose(Xg.grad[lo:hi], Xr.grad[lo:hi], rtol=1e-10, atol=1e-12)) and bool((Xg.grad[:lo] == 0).all()))
        # symmetric Gram: equal slices of the pairs a <= b per rank, both mirror entries stored into every rank's copy
        for A in (12, 13):
            Gs = skb.distributed.compute_Gram_sharded(sk, X[:A], sym=True)
            ref = sk.compute_Gram(X[:A], X[:A], sym=True)
            ok.append(bool(torch.allclose(Gs, ref, rtol=1e-13, atol=0)) and bool(torch.equal(Gs, Gs.t())))
            ok.append(skb.distributed.last_gather == "peers_sym")
        # ... and the block-cyclic tiles + one all-reduce when the peer path is switched off
        skb.distributed._PeerGram.disabled = True
        Gs = skb.distributed.compute_Gram_sharded(sk, X, sym=True)
        ok.append(bool(torch.allclose(Gs, sk.compute_Gram(X, X, sym=True), rtol=1e-13, atol=0)) and skb.distributed.last_gather == "all_reduce_sym")
        skb.distributed._PeerGram.disabled = False
        q.put((rank, all(ok), path))
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
    res = [q.get(timeout=5) for _ in range(2)]
    assert {r[0]: r[1] for r in res} == {0: True, 1: True}, res
    print("gather path on NCCL:", res[0][2])
