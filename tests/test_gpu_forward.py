"""GPU parity tests of the forward path (run on the B200 box: `pytest -m gpu`).

CUDA path (through the C ABI, via the reference-shaped Python API) vs
  * the committed golden vectors (outputs of the unmodified reference), tolerance
    |G - G_ref| <= 1e-10 (|G_ref| + 1)   (BASELINE.json north_star);
  * the oracle on seeded inputs (shapes the oracle finishes in seconds);
  * BITWISE for the exact-arithmetic entry points fed identical static matrices / increments;
  * size-independent properties at the full BASELINE sizes (symmetry, sym flag, batch == diag(Gram),
    row-block invariance, constant path == 1).
"""
import numpy as np
import pytest
import torch

from tests._util import FWD_TOL, fwd_err, golden_names, load_golden, make_paths, static_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def skb():
    import sigkernel_b200
    return sigkernel_b200


@pytest.fixture(scope="module")
def O():
    from oracle import sigkernel_oracle
    return sigkernel_oracle


def _static(mod, meta):
    return static_of(mod, meta)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


FWD_OPS = ("kernel", "gram", "gram_sym", "mmd", "distance", "scoring")


@pytest.mark.parametrize("name", golden_names(FWD_OPS))
def test_forward_matches_reference_golden(skb, name):
    meta, z = load_golden(name)
    X, Y = _dev(z["X"]), _dev(z["Y"])
    sk = skb.SigKernel(_static(skb, meta), meta["dyadic_order"], _naive_solver=meta["naive"])
    op = meta["op"]
    if op == "kernel":
        assert fwd_err(sk.compute_kernel(X, Y).cpu().numpy(), z["K"]) <= FWD_TOL
    elif op == "gram":
        assert fwd_err(sk.compute_Gram(X, Y).cpu().numpy(), z["G"]) <= FWD_TOL
    elif op == "gram_sym":
        assert fwd_err(sk.compute_Gram(X, X, sym=True).cpu().numpy(), z["G"]) <= FWD_TOL
    elif op == "mmd":
        assert fwd_err(sk.compute_mmd(X, Y).cpu().numpy(), z["mmd"]) <= FWD_TOL
    elif op == "distance":
        assert fwd_err(sk.compute_distance(X, Y).cpu().numpy(), z["dist"]) <= FWD_TOL
    elif op == "scoring":
        assert fwd_err(sk.compute_scoring_rule(X, Y).cpu().numpy(), z["score"]) <= FWD_TOL


@pytest.mark.parametrize("name", golden_names(("gram_bwd", "gram_sym_bwd", "kernel_bwd")))
def test_forward_values_of_backward_goldens(skb, name):
    meta, z = load_golden(name)
    X, Y = _dev(z["X"]), _dev(z["Y"])
    sk = skb.SigKernel(_static(skb, meta), meta["dyadic_order"], _naive_solver=meta["naive"])
    if meta["op"] == "kernel_bwd":
        assert fwd_err(sk.compute_kernel(X, Y).cpu().numpy(), z["K"]) <= FWD_TOL
    else:
        sym = meta["op"] == "gram_sym_bwd"
        assert fwd_err(sk.compute_Gram(X, X if sym else Y, sym=sym).cpu().numpy(), z["G"]) <= FWD_TOL


SHAPES = [
    # A, B, M, N, D, d, kind
    (3, 4, 2, 2, 1, 0, "rand"),
    (3, 4, 2, 7, 2, 3, "rand"),
    (2, 3, 33, 17, 3, 1, "rand"),
    (2, 2, 64, 64, 5, 2, "rand"),
    (2, 2, 65, 40, 2, 1, "bm"),
    (1, 3, 100, 90, 4, 0, "bm"),
    (2, 2, 128, 128, 8, 2, "rand"),
    (1, 2, 200, 30, 2, 1, "bm"),
    (5, 1, 9, 12, 6, 4, "randn"),
    (2, 2, 7, 7, 2, 5, "bm"),
]


@pytest.mark.parametrize("A,B,M,N,D,d,kind", SHAPES)
@pytest.mark.parametrize("static", ["rbf", "linear"])
@pytest.mark.parametrize("naive", [False, True])
def test_gram_vs_oracle(skb, O, A, B, M, N, D, d, kind, static, naive):
    X = make_paths(kind, 100 + M, (A, M, D))
    Y = make_paths(kind, 200 + N, (B, N, D))
    if kind == "randn":
        X, Y = 0.3 * X, 0.3 * Y
    ref = O.compute_Gram(X, Y, O.RBFKernel(0.7) if static == "rbf" else O.LinearKernel(), d, naive=naive)
    sk = skb.SigKernel(skb.RBFKernel(0.7) if static == "rbf" else skb.LinearKernel(), d, _naive_solver=naive)
    got = sk.compute_Gram(X.cuda(), Y.cuda())
    assert got.shape == (A, B) and got.dtype == torch.float64 and got.is_cuda
    assert fwd_err(got.cpu().numpy(), ref.numpy()) <= FWD_TOL


@pytest.mark.parametrize("scale", [1.0, 0.5])
def test_batch_kernel_vs_oracle_linear_scale(skb, O, scale):
    X, Y = make_paths("bm", 1, (6, 20, 3)), make_paths("bm", 2, (6, 15, 3))
    ref = O.compute_kernel(X, Y, O.LinearKernel(scale), 2)
    got = skb.SigKernel(skb.LinearKernel(scale), 2).compute_kernel(X.cuda(), Y.cuda())
    assert got.shape == (6,)
    assert fwd_err(got.cpu().numpy(), ref.numpy()) <= FWD_TOL
    # Gram_matrix ignores `scale` in the reference (static_kernels.py:33) -- preserved
    refg = O.compute_Gram(X, Y, O.LinearKernel(scale), 1)
    gotg = skb.SigKernel(skb.LinearKernel(scale), 1).compute_Gram(X.cuda(), Y.cuda())
    assert fwd_err(gotg.cpu().numpy(), refg.numpy()) <= FWD_TOL


@pytest.mark.parametrize("naive", [False, True])
@pytest.mark.parametrize("shape", [(3, 4, 17, 23), (2, 2, 1, 1), (1, 3, 200, 5), (2, 1, 64, 255)])
def test_solve_increments_bitwise(skb, O, naive, shape):
    """Operator-level entry point in exact arithmetic == compiled C restatement of cython_backend, bit for bit."""
    rng = np.random.default_rng(0)
    inc = rng.uniform(-0.3, 0.3, size=shape)
    ref = O.solve_gram(inc, False, naive)[:, :, -1, -1]
    got = skb.ops.solve_increments(_dev(inc), naive=naive, exact=True).cpu().numpy()
    assert np.array_equal(got, ref)
    fast = skb.ops.solve_increments(_dev(inc), naive=naive, exact=False).cpu().numpy()
    assert fwd_err(fast, ref) <= 1e-11


@pytest.mark.parametrize("d", [0, 1, 2, 3])
@pytest.mark.parametrize("pairs", ["gram", "batch"])
def test_from_static_exact_bitwise(skb, O, d, pairs):
    """Plugin path in exact arithmetic: identical coarse static matrix in => identical bits out."""
    X, Y = make_paths("rand", 5, (3, 21, 3)), make_paths("rand", 6, (3, 14, 3))
    sk = O.RBFKernel(0.5)
    if pairs == "gram":
        Ks = sk.Gram_matrix(X, Y)
        ref = torch.from_numpy(O.solve_gram(O.increments(Ks, d).numpy()))[:, :, -1, -1]
    else:
        Ks = sk.batch_kernel(X, Y)
        ref = torch.from_numpy(O.solve_batch(O.increments(Ks, d).numpy()))[:, -1, -1]
    got = skb.ops.sigkernel_forward_from_static(Ks.cuda(), d, pairs, exact=True).cpu()
    assert torch.equal(got, ref)


def test_plugin_kernel_goes_through_from_static(skb, O):
    """A user-defined static kernel (no fused_spec) must work and match the oracle fed the same object."""
    class Poly:
        def batch_kernel(self, X, Y):
            return (1. + torch.bmm(X, Y.transpose(1, 2))) ** 2

        def Gram_matrix(self, X, Y):
            return (1. + torch.einsum('ipk,jqk->ijpq', X, Y)) ** 2

    X, Y = make_paths("bm", 7, (3, 12, 2)), make_paths("bm", 8, (4, 9, 2))
    ref = O.compute_Gram(X, Y, Poly(), 1)
    got = skb.SigKernel(Poly(), 1).compute_Gram(X.cuda(), Y.cuda())
    assert fwd_err(got.cpu().numpy(), ref.numpy()) <= FWD_TOL
    refk = O.compute_kernel(X, Y[:3, :, :], Poly(), 2)
    gotk = skb.SigKernel(Poly(), 2).compute_kernel(X.cuda(), Y[:3].cuda())
    assert fwd_err(gotk.cpu().numpy(), refk.numpy()) <= FWD_TOL


def test_subclass_of_builtin_is_not_fused(skb):
    class MyRBF(skb.RBFKernel):
        def Gram_matrix(self, X, Y):
            return 2. * super().Gram_matrix(X, Y)
    assert MyRBF(1.0).fused_spec(True) is None


def test_float32_io(skb, O):
    X, Y = make_paths("bm", 9, (3, 16, 2), torch.float32), make_paths("bm", 10, (2, 16, 2), torch.float32)
    got = skb.SigKernel(skb.RBFKernel(1.0), 1).compute_Gram(X.cuda(), Y.cuda())
    assert got.dtype == torch.float32
    ref = O.compute_Gram(X.double(), Y.double(), O.RBFKernel(1.0), 1)
    assert fwd_err(got.double().cpu().numpy(), ref.numpy()) <= 1e-6


def test_cpu_tensors_fail_loudly(skb):
    X = make_paths("rand", 0, (2, 5, 2))
    with pytest.raises(skb.SigKernelB200Error):
        skb.SigKernel(skb.RBFKernel(1.0), 0).compute_Gram(X, X)


def test_error_codes(skb):
    lib = skb._lib.lib
    assert lib.skb_sigkernel_fwd(None, None, 0, 2, 2, 1, 4, 2, 0, 1, 1.0, 0, 0, 0, None, None, 0, None) == -1
    assert lib.skb_sigkernel_fwd(None, None, 0, 2, 3, 4, 4, 2, 0, 1, 1.0, 0, 1, 0, None, None, 0, None) == -1
    assert lib.skb_sigkernel_fwd(None, None, 0, 2, 2, 4, 4, 2, 0, 7, 1.0, 0, 0, 0, None, None, 0, None) == -2
    assert lib.skb_sigkernel_fwd(None, None, 0, 2, 2, 4, 4, 2, 0, 1, 1.0, 0, 0, 0, None, None, 0, None) == -6


# ---- properties at the full BASELINE sizes ----------------------------------------------------------
def test_cfg3_full_size_properties(skb):
    """128x128, len 64, dim 5, dyadic 2, RBF: the golden block is the top-left corner of the full Gram
    (same seeded tensors), the Gram of (X, X) is symmetric, sym=True agrees, row blocks are invariant."""
    meta, z = load_golden("cfg3_gram_rbf")
    g = torch.Generator().manual_seed(0)
    X = torch.rand((128, 64, 5), dtype=torch.float64, generator=g).cuda()
    Y = torch.rand((128, 64, 5), dtype=torch.float64, generator=g).cuda()
    sk = skb.SigKernel(skb.RBFKernel(0.5), 2)
    G = sk.compute_Gram(X, Y)
    n = z["G"].shape[0]
    assert np.array_equal(X[:n].cpu().numpy(), z["X"])
    assert fwd_err(G[:n, :n].cpu().numpy(), z["G"]) <= FWD_TOL
    Gxx = sk.compute_Gram(X, X)
    assert fwd_err(Gxx.cpu().numpy(), Gxx.T.cpu().numpy()) <= 1e-12
    assert fwd_err(sk.compute_Gram(X, X, sym=True).cpu().numpy(), Gxx.cpu().numpy()) <= 1e-12
    # pairs are independent: a row block alone gives the same values -- bit for bit when the same kernel serves both
    # batches (the default plan may send the smaller batch to fwd5_kernel and the full one to the tile kernel, whose
    # exp tables differ in the last bit)
    assert fwd_err(sk.compute_Gram(X[32:96], Y).cpu().numpy(), G[32:96].cpu().numpy()) <= 1e-12
    lib = skb._lib.lib
    for mode in (0, 1):
        lib.skb_set_tile_mode(mode)
        try:
            Gm = sk.compute_Gram(X, Y)
            assert torch.equal(sk.compute_Gram(X[32:96], Y), Gm[32:96])
            assert fwd_err(Gm.cpu().numpy(), G.cpu().numpy()) <= 1e-12
        finally:
            lib.skb_set_tile_mode(-1)
    assert fwd_err(sk.compute_kernel(X, Y).cpu().numpy(), torch.diag(G).cpu().numpy()) <= 1e-12
    # deterministic: same launch twice gives the same bits
    assert torch.equal(sk.compute_Gram(X, Y), G)


def test_cfg2_full_size_vs_golden_corner(skb):
    meta, z = load_golden("cfg2_gram_rbf")
    g = torch.Generator().manual_seed(0)
    X = torch.rand((64, 32, 3), dtype=torch.float64, generator=g).cuda()
    Y = torch.rand((64, 32, 3), dtype=torch.float64, generator=g).cuda()
    G = skb.SigKernel(skb.RBFKernel(0.5), 1).compute_Gram(X, Y)
    n = z["G"].shape[0]
    assert fwd_err(G[:n, :n].cpu().numpy(), z["G"]) <= FWD_TOL


def test_constant_path_gives_one(skb):
    X = (torch.ones(3, 40, 3, dtype=torch.float64) * 0.37).cuda()
    Y = make_paths("randn", 3, (2, 33, 3)).cuda()
    for k in (skb.LinearKernel(), skb.RBFKernel(0.7)):
        G = skb.SigKernel(k, 2).compute_Gram(X, Y)
        assert fwd_err(G.cpu().numpy(), np.ones((3, 2))) <= 1e-13


def test_many_pairs_stream_through_few_warps(skb, O):
    """Force 1 warp per SM so that every warp streams several pairs back to back (job switching)."""
    X, Y = make_paths("rand", 31, (40, 9, 2)), make_paths("rand", 32, (37, 6, 2))
    ref = O.compute_Gram(X, Y, O.RBFKernel(0.5), 1)
    skb._lib.lib.skb_set_warps_per_sm(1)
    try:
        got = skb.SigKernel(skb.RBFKernel(0.5), 1).compute_Gram(X.cuda(), Y.cuda())
    finally:
        skb._lib.lib.skb_set_warps_per_sm(0)
    assert fwd_err(got.cpu().numpy(), ref.numpy()) <= FWD_TOL


# ---- shapes outside the register-resident kernels: generic row-band fallback ------------------------
GENERIC_SHAPES = [
    # A, B, M, N, D, d   (ceil(M/32) * 2^d > 32, or M > 256)
    (2, 2, 300, 20, 2, 0),
    (1, 2, 300, 9, 3, 1),
    (2, 1, 10, 12, 2, 6),
    (1, 1, 1000, 6, 2, 0),      # the reference's own limit is max(MM, NN) < 1024
    (1, 2, 70, 150, 2, 3),
    (2, 2, 40, 33, 4, 5),
]


@pytest.mark.parametrize("A,B,M,N,D,d", GENERIC_SHAPES)
@pytest.mark.parametrize("static", ["rbf", "linear"])
def test_generic_shapes_vs_oracle(skb, O, A, B, M, N, D, d, static):
    X = make_paths("bm", 500 + M, (A, M, D))
    Y = make_paths("bm", 600 + N, (B, N, D))
    ref = O.compute_Gram(X, Y, O.RBFKernel(0.9) if static == "rbf" else O.LinearKernel(), d)
    sk = skb.SigKernel(skb.RBFKernel(0.9) if static == "rbf" else skb.LinearKernel(), d)
    got = sk.compute_Gram(X.cuda(), Y.cuda())
    assert fwd_err(got.cpu().numpy(), ref.numpy()) <= FWD_TOL
    if A == B:
        refk = O.compute_kernel(X, Y, O.RBFKernel(0.9) if static == "rbf" else O.LinearKernel(), d)
        assert fwd_err(sk.compute_kernel(X.cuda(), Y.cuda()).cpu().numpy(), refk.numpy()) <= FWD_TOL
        if M == N:
            assert fwd_err(sk.compute_Gram(X.cuda(), X.cuda(), sym=True).cpu().numpy(),
                           O.compute_Gram(X, X, O.RBFKernel(0.9) if static == "rbf" else O.LinearKernel(), d).numpy()) <= FWD_TOL


def test_generic_exact_bitwise(skb, O):
    """The row-band fallback in exact arithmetic is still bit-identical to the reference solver."""
    rng = np.random.default_rng(3)
    inc = rng.uniform(-0.2, 0.2, size=(2, 700, 9))
    ref = O.solve_batch(inc)[:, -1, -1]
    assert np.array_equal(skb.ops.solve_increments(_dev(inc), exact=True).cpu().numpy(), ref)
    X, Y = make_paths("rand", 5, (2, 290, 2)), make_paths("rand", 6, (2, 7, 2))
    Ks = O.RBFKernel(0.5).Gram_matrix(X, Y)
    ref = torch.from_numpy(O.solve_gram(O.increments(Ks, 1).numpy()))[:, :, -1, -1]
    assert torch.equal(skb.ops.sigkernel_forward_from_static(Ks.cuda(), 1, "gram", exact=True).cpu(), ref)


# ---- fwd5_kernel (skb_fwd5.cuh): job ring, virtual start, 1 / 2 / 4 warps per pair ---------------------
FWD5_SHAPES = [
    # A, B, M, N, D, d    what it stresses
    (37, 41, 9, 4, 2, 1),      # N = 4: production a whole pair ahead, a ring entry every 4 steps, lanes >= N start virtual several times
    (23, 19, 12, 5, 3, 2),     # N = 5, wrap column 0
    (3, 2, 64, 64, 5, 2),      # the headline strip (RC=2, d=2, Dp=6)
    (5, 4, 33, 6, 9, 0),       # Dp = 10, RC = 2, d = 0
    (3, 3, 100, 11, 8, 1),     # RC = 4 at one warp would be 8 rows: 2 warps per pair (RC = 2)
    (2, 3, 128, 128, 8, 2),    # cfg5's strip: 2 warps per pair
    (2, 2, 250, 9, 3, 2),      # 4 warps per pair (RC = 2)
    (2, 2, 70, 7, 1, 3),       # d = 3: RC = 1 with 4 warps per pair
    (6, 6, 31, 31, 4, 3),      # RC = 1, d = 3, one warp
]


@pytest.mark.parametrize("A,B,M,N,D,d", FWD5_SHAPES)
@pytest.mark.parametrize("static", ["rbf", "linear"])
def test_fwd5_shapes_vs_oracle(skb, O, A, B, M, N, D, d, static):
    X = make_paths("bm", 700 + M, (A, M, D))
    Y = make_paths("bm", 800 + N, (B, N, D))
    ok = O.RBFKernel(0.8) if static == "rbf" else O.LinearKernel()
    sk = skb.SigKernel(skb.RBFKernel(0.8) if static == "rbf" else skb.LinearKernel(), d)
    for wpsm in (0, 4):        # default residency, and few resident warps so that every stream runs many pairs
        skb._lib.lib.skb_set_warps_per_sm(wpsm)
        try:
            got = sk.compute_Gram(X.cuda(), Y.cuda())
        finally:
            skb._lib.lib.skb_set_warps_per_sm(0)
        assert fwd_err(got.cpu().numpy(), O.compute_Gram(X, Y, ok, d).numpy()) <= FWD_TOL
    if A == B:
        assert fwd_err(sk.compute_kernel(X.cuda(), Y.cuda()).cpu().numpy(), O.compute_kernel(X, Y, ok, d).numpy()) <= FWD_TOL
        if M == N:
            Xc = X.cuda()
            Gs = sk.compute_Gram(Xc, Xc, sym=True)
            assert torch.equal(Gs, Gs.T)
            assert fwd_err(Gs.cpu().numpy(), O.compute_Gram(X, X, ok, d).numpy()) <= FWD_TOL


def test_fwd5_long_stream_single_block(skb, O):
    """One resident block streams 1500 pairs of 4-point paths: exercises the ring far beyond its depth."""
    X, Y = make_paths("rand", 91, (30, 5, 2)), make_paths("rand", 92, (50, 4, 2))
    ref = O.compute_Gram(X, Y, O.RBFKernel(0.5), 2)
    import ctypes
    lib = skb._lib.lib
    Xc, Yc = X.cuda(), Y.cuda()
    out = torch.empty(30 * 50, dtype=torch.float64, device="cuda")
    nbytes = lib.skb_fwd_workspace_bytes(30, 50, 5, 4, 2, 2, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    # warps_per_sm = 1 still gives 148 blocks; 1500 pairs / 148 blocks ~ 10 pairs per stream
    lib.skb_set_warps_per_sm(1)
    try:
        skb._lib.check(lib.skb_sigkernel_fwd(Xc.data_ptr(), Yc.data_ptr(), 0, 30, 50, 5, 4, 2, 2, 1, ctypes.c_double(0.5), 0, 0, 0,
                                             out.data_ptr(), ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream))
    finally:
        lib.skb_set_warps_per_sm(0)
    assert fwd_err(out.view(30, 50).cpu().numpy(), ref.numpy()) <= FWD_TOL


def test_overflowing_pair_does_not_leak_into_its_stream(skb, O):
    """A pair whose PDE solution overflows fp64 returns inf / NaN (as the reference does) and the pairs that follow
    it in the same warp stream are unaffected: every per-pair state is re-armed by assignment, not arithmetic."""
    X = make_paths("bm", 61, (5, 20, 3))
    Y = make_paths("bm", 62, (30, 18, 3))
    X[2] *= 1e90                                   # increments ~1e90: the grid of every pair (2, b) overflows
    ref = O.compute_Gram(X, Y, O.LinearKernel(), 2).numpy()
    bad = ~np.isfinite(ref)
    assert bad[2].all() and not bad[[0, 1, 3, 4]].any()
    for wpsm in (0, 1):                            # 1: few streams, every one of them runs through overflowed pairs
        skb._lib.lib.skb_set_warps_per_sm(wpsm)
        try:
            got = skb.SigKernel(skb.LinearKernel(), 2).compute_Gram(X.cuda(), Y.cuda()).cpu().numpy()
        finally:
            skb._lib.lib.skb_set_warps_per_sm(0)
        assert (~np.isfinite(got[2])).all()
        assert fwd_err(got[~bad], ref[~bad]) <= FWD_TOL


@pytest.mark.parametrize("pairs", ["gram", "sym"])
def test_forward_range_slices_assemble_the_full_matrix(skb, pairs):
    """skb_sigkernel_fwd_range: slices of the pair enumeration (one rank's share of a sharded Gram matrix) written into a
    local matrix, and -- with one "peer", this device -- through the peer-store path with the in-kernel rank barrier."""
    X = make_paths("rand", 7, (11, 20, 3)).cuda()
    Y = X if pairs == "sym" else make_paths("rand", 8, (9, 17, 3)).cuda()
    sk = skb.SigKernel(skb.RBFKernel(0.5), 1)
    ref = sk.compute_Gram(X, Y, sym=pairs == "sym")
    A, B = X.shape[0], Y.shape[0]
    total = skb.ops.n_jobs(A, B, pairs)
    out = torch.full((A, B), float("nan"), dtype=torch.float64, device="cuda")
    cuts = [0, total // 3, total // 3, (2 * total) // 3 + 1, total]          # includes an empty slice
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        assert skb.ops.sigkernel_forward_range(X, Y, "rbf", 0.5, 1, lo, hi, pairs, out=out)
    assert torch.equal(out, ref)
    # peer stores into "every rank's copy" (here: one) + the kernel's own barrier: slot 0 of the signal array reaches the epoch
    buf = torch.full((A, B), float("nan"), dtype=torch.float64, device="cuda")
    sig = torch.zeros(8, dtype=torch.int64, device="cuda")
    for epoch, (lo, hi) in enumerate(((0, total // 2), (total // 2, total), (total, total)), start=1):
        assert skb.ops.sigkernel_forward_range(X, Y, "rbf", 0.5, 1, lo, hi, pairs, peer_ptrs=[buf.data_ptr()],
                                               signal=([sig.data_ptr()], 0, epoch))
        torch.cuda.synchronize()
        assert int(sig[0]) == epoch
    assert torch.equal(buf, ref)


def test_static_gram_matches_the_plugin_kernels(skb):
    """skb_static_gram (one CUDA pass) against the torch implementation behind the plugin interface."""
    X, Y = make_paths("rand", 3, (4, 9, 3)).cuda(), make_paths("rand", 4, (5, 7, 3)).cuda()
    for k, kind, par in ((skb.RBFKernel(0.7), "rbf", 0.7), (skb.LinearKernel(), "linear", 1.0)):
        K = skb.ops.static_gram(X, Y, kind, par, "gram")
        assert K.shape == (4, 5, 9, 7)
        assert fwd_err(K.cpu().numpy(), k.Gram_matrix(X, Y).cpu().numpy()) <= 1e-14
    Kb = skb.ops.static_gram(X, Y[:4, :, :], "rbf", 0.7, "batch")
    assert fwd_err(Kb.cpu().numpy(), skb.RBFKernel(0.7).batch_kernel(X, Y[:4]).cpu().numpy()) <= 1e-14
