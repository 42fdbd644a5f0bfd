"""Multi-GPU Gram matrix: the (a, b) pairs are independent PDE solves, so the Gram axis shards with
no data-path exchange -- contiguous row blocks of X per rank, Y replicated, one all-gather of the
(rows_r, B) fp64 blocks to reassemble G.  The reference has no multi-device support.

One process per GPU (torchrun); `torch.distributed` must be initialised by the caller.  The solve is
injected as a callable so that the sharding logic is testable on CPU with the gloo backend.
"""
import torch
import torch.distributed as dist


def row_block(n_rows, rank, world):
    """Contiguous block [lo, hi) of rank `rank`; the first n_rows % world ranks get one more row."""
    q, r = divmod(n_rows, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


class _GatherRows(torch.autograd.Function):
    """All-gather of equally sized row blocks that autograd can cross: the backward of rank r is simply rows
    [lo_r, hi_r) of the upstream gradient -- no collective, PROVIDED every rank evaluates the same loss on the
    gathered matrix (the usual data-parallel situation: G is replicated, so is d loss / d G)."""

    @staticmethod
    def forward(ctx, block, lo, hi, n_rows, group):
        ctx.lo, ctx.hi = lo, hi
        G = torch.empty((n_rows, block.shape[1]), dtype=block.dtype, device=block.device)
        dist.all_gather_into_tensor(G, block.contiguous(), group=group)
        return G

    @staticmethod
    def backward(ctx, grad_G):
        return grad_G[ctx.lo:ctx.hi], None, None, None, None


def sharded_gram(X, Y, gram_fn, group=None, gather=True):
    """G = gram_fn(X, Y) computed as row blocks: rank r solves gram_fn(X[lo_r:hi_r], Y).

    X (A, M, D) and Y (B, N, D) are the FULL inputs on every rank (paths are small: cfg5 is 4 MB each).
    Returns the full (A, B) matrix on every rank (gather=True) or this rank's (hi-lo, B) block.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    A, B = X.shape[0], Y.shape[0]
    lo, hi = row_block(A, rank, world)
    if hi > lo:
        block = gram_fn(X[lo:hi], Y)
    else:
        block = torch.empty((0, B), dtype=X.dtype, device=X.device)
    if not gather:
        return block
    if A % world == 0:
        if block.requires_grad:
            # gradients w.r.t. X flow back into this rank's rows of X only (each rank owns the complete rows a
            # of grad_points[a, :, :, :]); all-reduce X.grad afterwards if every rank needs all rows
            return _GatherRows.apply(block, lo, hi, A, group)
        G = torch.empty((A, B), dtype=block.dtype, device=block.device)
        dist.all_gather_into_tensor(G, block.contiguous(), group=group)
        return G
    # ragged split: pad every block to the largest one, gather, drop the padding
    rows_max = -(-A // world)
    padded = torch.zeros((rows_max, B), dtype=block.dtype, device=block.device)
    padded[:hi - lo] = block
    buf = torch.empty((world * rows_max, B), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    parts = []
    for r in range(world):
        l, h = row_block(A, r, world)
        parts.append(buf[r * rows_max: r * rows_max + (h - l)])
    return torch.cat(parts, dim=0)


def compute_Gram_sharded(sig_kernel, X, Y, group=None, gather=True):
    """`SigKernel.compute_Gram(X, Y)` sharded over the ranks of `group`.  Differentiable w.r.t. X when the batch
    divides evenly: after `loss(G).backward()` rank r holds d loss / d X in rows [lo_r, hi_r) of `X.grad` (zeros
    elsewhere); `all_reduce_grad_rows(X.grad)` replicates the full gradient if it is needed everywhere."""
    return sharded_gram(X, Y, lambda x, y: sig_kernel.compute_Gram(x, y, sym=False), group, gather)


def all_reduce_grad_rows(grad, group=None):
    """Sum the per-rank row-sparse gradients of a sharded backward so that every rank holds all rows."""
    dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    return grad
