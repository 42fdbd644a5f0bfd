"""Multi-GPU Gram matrix: the (a, b) pairs are independent PDE solves, so the Gram axis shards with
no data-path exchange -- contiguous row blocks of X per rank, Y replicated, one all-gather of the
(rows_r, B) fp64 blocks to reassemble G.  The reference has no multi-device support.

One process per GPU (torchrun); `torch.distributed` must be initialised by the caller.  The solve is
injected as a callable so that the sharding logic is testable on CPU with the gloo backend.

On CUDA with the NCCL backend the gather needs no collective at all (`_PeerGram`): G lives in symmetric memory
(`torch.distributed._symmetric_memory`: every rank maps every rank's buffer over NVLink / NVSwitch), the solver kernel
stores each k(X_a, Y_b) -- 8 bytes -- straight into its place in EVERY rank's copy while it runs
(skb_sigkernel_fwd_peers), and one barrier across the ranks follows.  Round 1 measured the NCCL all-gather of the 128 KB
blocks at 40-50 us behind a 0.35 ms solve (88 % weak-scaling efficiency); the stores overlap the solve completely.
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


def sym_tiles(n_rows, world):
    """Block-cyclic assignment of the upper triangle of a symmetric (n_rows, n_rows) matrix: 2 * world row blocks, the
    tiles (I, J), I <= J, ordered by diagonal offset and dealt to the ranks round-robin -- every rank gets two diagonal
    tiles (half the work of a full one) and 2 * world - 1 off-diagonal ones.  (Contiguous row blocks of the triangle would
    give the first rank about twice the work of the last.)  Returns (row bounds of the blocks, [(I, J, rank), ...])."""
    nb = max(1, min(2 * world, n_rows))
    bounds = [row_block(n_rows, i, nb) for i in range(nb)]
    tiles = []
    for off in range(nb):
        for i in range(nb - off):
            tiles.append((i, i + off, len(tiles) % world))
    return bounds, tiles


def sharded_gram_sym(X, gram_fn, group=None):
    """G = gram_fn(X, X, sym=True) with the unordered pairs sharded: each rank solves its tiles of the upper triangle
    (gram_fn(x_I, x_I, True) on the diagonal, gram_fn(x_I, x_J, False) mirrored off it) into a zero matrix; one all-reduce
    (sum) assembles the full symmetric matrix on every rank.  No gradient flows through it."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    A = X.shape[0]
    bounds, tiles = sym_tiles(A, world)
    G = None
    for I, J, r in tiles:
        if r != rank:
            continue
        (li, hi), (lj, hj) = bounds[I], bounds[J]
        if hi == li or hj == lj:
            continue
        blk = gram_fn(X[li:hi], X[lj:hj], I == J).detach()
        if G is None:
            G = torch.zeros((A, A), dtype=blk.dtype, device=blk.device)
        G[li:hi, lj:hj] = blk
        if I != J:
            G[lj:hj, li:hi] = blk.transpose(0, 1)
    if G is None:
        G = torch.zeros((A, A), dtype=X.dtype if X.dtype == torch.float64 else torch.float64, device=X.device)
    dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)
    return G


# how the ranks synchronise after a peer-store solve: "handle" = a symmetric-memory barrier kernel behind the solve,
# "kernel" = the solver kernel's last block signals and waits itself (skb_sigkernel_fwd_range).  Measured on B200 (DESIGN.md
# 5): no difference beyond run-to-run noise at 2 ranks (0.37 ms per step either way), the separate barrier is faster at 8
# (0.372 vs 0.385 ms) -- the default.
RANK_BARRIER = __import__("os").environ.get("SKB_RANK_BARRIER", "handle")

last_gather = None     # "peers" / "all_gather": which path the last compute_Gram_sharded(gather=True) call took (diagnostic)


class _PeerGram:
    """Symmetric-memory buffers for the collective-free gather, cached per (rows, columns, device, group): two copies
    of G used alternately, so that a rank that is one call ahead never overwrites a result a peer may still be reading
    (a rank passes the barrier of call k only after every rank has finished the solve of call k, and starts writing call
    k + 2 into the buffer of call k only after the barrier of call k + 1)."""
    _cache = {}
    disabled = False

    def __init__(self, n_rows, n_cols, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        self.bufs, self.hdls = [], []
        for _ in range(2):
            t = symm_mem.empty((n_rows, n_cols), dtype=torch.float64, device=device)
            self.hdls.append(symm_mem.rendezvous(t, grp))
            self.bufs.append(t)
        self.turn = 0
        # signal slots of the in-kernel rank barrier (skb_sigkernel_fwd_range): one 64-bit slot per rank, epoch-valued
        self.sig = symm_mem.empty((64,), dtype=torch.int64, device=device)
        self.sig.zero_()
        self.sig_hdl = symm_mem.rendezvous(self.sig, grp)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)                      # every rank's slots are zero before anybody signals
        self.epoch = 0
        self.rank = dist.get_rank(group)

    @classmethod
    def get(cls, n_rows, n_cols, device, group):
        key = (n_rows, n_cols, str(device), id(group))
        if key not in cls._cache:
            cls._cache[key] = cls(n_rows, n_cols, device, group)
        return cls._cache[key]


def _gram_into_peers(sig_kernel, X, Y, lo, hi, group):
    """Rows [lo, hi) of Gram(X, Y) written into every rank's symmetric-memory copy of G by the solver kernel itself.
    Returns the local copy (valid until the next-but-one sharded call of the same shape) or None if this path does not
    apply (CPU / gloo, plugin static kernels, gradients, shapes outside fwd5_kernel, no symmetric memory)."""
    if _PeerGram.disabled or not X.is_cuda or dist.get_backend(group) != "nccl":
        return None
    if torch.is_grad_enabled() and (X.requires_grad or Y.requires_grad):
        return None
    spec = getattr(sig_kernel.static_kernel, "fused_spec", None)
    spec = spec(True) if spec is not None else None
    if spec is None or spec[2] is not None or X.dim() != 3 or X.dtype != torch.float64:
        return None
    from . import ops
    try:
        pg = _PeerGram.get(X.shape[0], Y.shape[0], X.device, group)
    except Exception:                                  # symmetric memory unavailable on this system: use the all-gather
        _PeerGram.disabled = True
        return None
    if ops.lib.skb_forward_plan(X.shape[1], Y.shape[1], X.shape[2], int(sig_kernel.dyadic_order), ops._STATIC[spec[0]],
                                ops._lib.SCHEME_S1 if sig_kernel._naive_solver else ops._lib.SCHEME_S2) < 4:
        return None                                    # (decided from the shape alone: every rank takes the same branch)
    k = pg.turn
    pg.turn ^= 1
    hdl, buf = pg.hdls[k], pg.bufs[k]
    pg.epoch += 1
    B = Y.shape[0]
    # rows [lo, hi) = jobs [lo B, hi B) of the GRAM enumeration; the kernel's last block is the barrier across the ranks
    in_kernel = RANK_BARRIER == "kernel"
    ops.sigkernel_forward_range(X, Y, spec[0], spec[1], sig_kernel.dyadic_order, lo * B, hi * B, "gram",
                                peer_ptrs=[int(q) for q in hdl.buffer_ptrs], naive=sig_kernel._naive_solver,
                                signal=([int(q) for q in pg.sig_hdl.buffer_ptrs], pg.rank, pg.epoch) if in_kernel else None)
    if not in_kernel:
        hdl.barrier(channel=0)
    return buf


def _gram_sym_into_peers(sig_kernel, X, group):
    """Gram(X, X): every rank solves an equal slice of the pairs a <= b (the SYM enumeration of skb_sigkernel_fwd_range) and
    the solver kernel stores each value -- and its mirror entry -- into every rank's symmetric-memory copy of G.  Returns the
    local copy or None if this path does not apply (see _gram_into_peers)."""
    if _PeerGram.disabled or not X.is_cuda or dist.get_backend(group) != "nccl":
        return None
    spec = getattr(sig_kernel.static_kernel, "fused_spec", None)
    spec = spec(True) if spec is not None else None
    if spec is None or spec[2] is not None or X.dim() != 3 or X.dtype != torch.float64:
        return None
    from . import ops
    A = X.shape[0]
    try:
        pg = _PeerGram.get(A, A, X.device, group)
    except Exception:                                  # symmetric memory unavailable on this system
        _PeerGram.disabled = True
        return None
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if ops.lib.skb_forward_plan(X.shape[1], X.shape[1], X.shape[2], int(sig_kernel.dyadic_order), ops._STATIC[spec[0]],
                                ops._lib.SCHEME_S1 if sig_kernel._naive_solver else ops._lib.SCHEME_S2) < 4:
        return None
    total = ops.n_jobs(A, A, "sym")
    lo, hi = row_block(total, rank, world)
    k = pg.turn
    pg.turn ^= 1
    hdl, buf = pg.hdls[k], pg.bufs[k]
    pg.epoch += 1
    in_kernel = RANK_BARRIER == "kernel"
    ops.sigkernel_forward_range(X, X, spec[0], spec[1], sig_kernel.dyadic_order, lo, hi, "sym",
                                peer_ptrs=[int(q) for q in hdl.buffer_ptrs], naive=sig_kernel._naive_solver,
                                signal=([int(q) for q in pg.sig_hdl.buffer_ptrs], pg.rank, pg.epoch) if in_kernel else None)
    if not in_kernel:
        hdl.barrier(channel=0)
    return buf


def compute_Gram_sharded(sig_kernel, X, Y=None, group=None, gather=True, sym=False):
    """`SigKernel.compute_Gram(X, Y, sym)` sharded over the ranks of `group`.  Differentiable w.r.t. X when the batch
    divides evenly: after `loss(G).backward()` rank r holds d loss / d X in rows [lo_r, hi_r) of `X.grad` (zeros
    elsewhere); `all_reduce_grad_rows(X.grad)` replicates the full gradient if it is needed everywhere.

    sym=True (Y is X): the unordered pairs are sharded instead of the rows -- equal slices of the pair enumeration with
    peer stores on CUDA / NCCL, block-cyclic tiles of the upper triangle and one all-reduce otherwise (`sym_tiles`) -- so
    every rank does 1 / world of HALF the square.  Forward values only: with gradients enabled and X requiring them the
    call takes the row-sharded path of the full square."""
    if sym:
        if Y is not None and Y is not X and not (Y.shape == X.shape and torch.equal(Y, X)):
            raise ValueError("sym=True needs Y = X")
        if gather and not (torch.is_grad_enabled() and X.requires_grad):
            G = _gram_sym_into_peers(sig_kernel, X, group)
            if G is not None:
                globals()["last_gather"] = "peers_sym"
                return G
            globals()["last_gather"] = "all_reduce_sym"
            return sharded_gram_sym(X, lambda x, y, s: sig_kernel.compute_Gram(x, y, sym=s), group)
        Y = X
    if gather and X.shape[0] % dist.get_world_size(group) == 0:
        lo, hi = row_block(X.shape[0], dist.get_rank(group), dist.get_world_size(group))
        G = _gram_into_peers(sig_kernel, X, Y, lo, hi, group)
        if G is not None:
            globals()["last_gather"] = "peers"
            return G
    globals()["last_gather"] = "all_gather"
    return sharded_gram(X, Y, lambda x, y: sig_kernel.compute_Gram(x, y, sym=False), group, gather)


def all_reduce_grad_rows(grad, group=None):
    """Sum the per-rank row-sparse gradients of a sharded backward so that every rank holds all rows."""
    dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    return grad
