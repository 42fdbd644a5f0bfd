"""world_size-2 gloo tests (CPU) of the Gram sharding logic (sigkernel_b200/distributed.py): row blocks
per rank, Y replicated, one all-gather -- with the solve injected (the oracle stands in for the kernel)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sigkernel_b200.distributed import row_block


def test_row_block_partition():
    for n in (1, 7, 8, 128, 513):
        for w in (1, 2, 3, 8):
            blocks = [row_block(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [h - l for l, h in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, A, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sigkernel_oracle as O
        from sigkernel_b200.distributed import sharded_gram
        g = torch.Generator().manual_seed(0)
        X = torch.rand((A, 6, 2), dtype=torch.float64, generator=g)
        Y = torch.rand((5, 7, 2), dtype=torch.float64, generator=g)
        fn = lambda x, y: O.compute_Gram(x, y, O.RBFKernel(0.5), 1)
        G = sharded_gram(X, Y, fn)
        blk = sharded_gram(X, Y, fn, gather=False)
        ref = fn(X, Y)
        lo, hi = row_block(A, rank, world)
        ok = torch.equal(G, ref) and torch.equal(blk, ref[lo:hi])
        out_q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("A", [6, 7])
def test_sharded_gram_world2_gloo(A):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, A, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}


def _worker_grad(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sigkernel_b200.distributed import all_reduce_grad_rows, sharded_gram
        g = torch.Generator().manual_seed(0)
        X = torch.rand((6, 5, 2), dtype=torch.float64, generator=g)
        Y = torch.rand((4, 7, 2), dtype=torch.float64, generator=g)
        w = torch.rand((6, 4), dtype=torch.float64, generator=g)
        fn = lambda x, y: torch.einsum('amd,bnd->ab', torch.tanh(x), y)      # any differentiable stand-in for the solve
        Xr = X.clone().requires_grad_(True)
        (fn(Xr, Y) * w).sum().backward()
        Xs = X.clone().requires_grad_(True)
        G = sharded_gram(Xs, Y, fn)
        (G * w).sum().backward()
        lo, hi = row_block(6, rank, world)
        own = torch.allclose(Xs.grad[lo:hi], Xr.grad[lo:hi], rtol=0, atol=1e-14)
        rest = torch.count_nonzero(Xs.grad[:lo]) == 0 and torch.count_nonzero(Xs.grad[hi:]) == 0
        full = all_reduce_grad_rows(Xs.grad.clone())
        out_q.put((rank, bool(own and rest and torch.allclose(full, Xr.grad, rtol=0, atol=1e-14) and torch.equal(G.detach(), fn(X, Y)))))
    finally:
        dist.destroy_process_group()


def test_sharded_gram_backward_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_grad, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(q.get(timeout=5) for _ in range(2)) == {0: True, 1: True}


def test_sym_tiles_cover_the_triangle_once_and_balance():
    from sigkernel_b200.distributed import sym_tiles
    for n in (1, 3, 7, 128, 513):
        for w in (1, 2, 3, 8):
            bounds, tiles = sym_tiles(n, w)
            nb = len(bounds)
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert sorted((i, j) for i, j, _ in tiles) == [(i, j) for i in range(nb) for j in range(i, nb)]
            if nb == 2 * w:
                cost = [sum(0.5 if i == j else 1.0 for i, j, r in tiles if r == k) for k in range(w)]
                assert max(cost) == min(cost) == 2 * w      # (2w)^2 / 2 tiles' worth of work, split evenly


def _worker_sym(rank, world, port, A, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sigkernel_oracle as O
        from sigkernel_b200.distributed import sharded_gram_sym
        g = torch.Generator().manual_seed(0)
        X = torch.rand((A, 6, 2), dtype=torch.float64, generator=g)
        calls = []

        def fn(x, y, s):
            calls.append((x.shape[0], y.shape[0], s))
            return O.compute_Gram(x, y, O.RBFKernel(0.5), 1, sym=s)
        G = sharded_gram_sym(X, fn)
        ref = O.compute_Gram(X, X, O.RBFKernel(0.5), 1, sym=True)
        ok = torch.allclose(G, ref, rtol=1e-14, atol=0) and torch.equal(G, G.t())
        # this rank solved its share only: the pairs it touched add up to about half the triangle
        pairs = sum(a * (a + 1) // 2 if s else a * b for a, b, s in calls)
        ok = ok and pairs <= (A * (A + 1) // 2) // world + 2 * A
        out_q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("A", [3, 8, 11])
def test_sharded_gram_sym_world2_gloo(A):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sym, args=(r, 2, port, A, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(q.get(timeout=5) for _ in range(2)) == {0: True, 1: True}
