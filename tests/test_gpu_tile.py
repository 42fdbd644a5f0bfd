"""GPU parity tests of the tile forward kernel (skb_tile.cuh: one pair per lane, one strip per warp, W-warp
pipelines, bands handed over through global memory), forced on through skb_set_tile_mode(1) for shapes the
default heuristic would leave to fwd5_kernel.  Oracle = checker only."""
import numpy as np
import pytest
import torch

from tests._util import FWD_TOL, fwd_err, load_golden, make_paths

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def skb():
    import sigkernel_b200
    return sigkernel_b200


@pytest.fixture(scope="module")
def O():
    from oracle import sigkernel_oracle
    return sigkernel_oracle


@pytest.fixture()
def tile_on(skb):
    skb._lib.lib.skb_set_tile_mode(1)
    yield
    skb._lib.lib.skb_set_tile_mode(-1)


SHAPES = [
    # A, B, M, N, D, d, kind        (strips of 16 fine rows, bands of 8 strips = 128 fine rows)
    (3, 4, 5, 4, 2, 2, "rand"),        # one strip, lanes mostly idle
    (33, 2, 9, 7, 3, 1, "rand"),       # two tiles per column, ragged second tile
    (40, 3, 33, 17, 5, 2, "rand"),     # exactly one band (128 fine rows)
    (5, 2, 34, 12, 3, 2, "rand"),      # 132 fine rows: two bands, the second with one live strip
    (64, 2, 64, 64, 5, 2, "rand"),     # the headline pair shape: 252 rows = two bands
    (2, 3, 100, 20, 8, 2, "bm"),       # 396 rows: four bands, x rows staged in shared memory (D + 1 = 10 wide)
    (3, 2, 70, 9, 2, 1, "bm"),         # dyadic order 1: 8 coarse rows per strip
    (2, 2, 20, 30, 3, 3, "bm"),        # dyadic order 3: 2 coarse rows per strip, 152 rows
    (1, 1, 300, 5, 1, 1, "bm"),        # 598 rows: five bands of one tile, every hand-off waits on the band above
]


@pytest.mark.parametrize("A,B,M,N,D,d,kind", SHAPES)
@pytest.mark.parametrize("static", ["rbf", "linear"])
def test_tile_gram_vs_oracle(skb, O, tile_on, A, B, M, N, D, d, kind, static):
    X, Y = make_paths(kind, 700 + M, (A, M, D)), make_paths(kind, 800 + N, (B, N, D))
    ok = O.RBFKernel(0.9) if static == "rbf" else O.LinearKernel()
    ref = O.compute_Gram(X, Y, ok, d).numpy()
    assert skb._lib.lib.skb_forward_plan(M, N, D, d, 1 if static == "rbf" else 0, 0) >= 0
    got = skb.ops.sigkernel_forward(X.cuda(), Y.cuda(), static, 0.9 if static == "rbf" else 1.0, d, "gram")
    assert fwd_err(got.cpu().numpy(), ref) <= FWD_TOL


def test_tile_batch_vs_oracle(skb, O, tile_on):
    X, Y = make_paths("rand", 901, (37, 20, 3)), make_paths("rand", 902, (37, 14, 3))
    ref = O.compute_kernel(X, Y, O.RBFKernel(0.5), 2).numpy()
    got = skb.ops.sigkernel_forward(X.cuda(), Y.cuda(), "rbf", 0.5, 2, "batch")
    assert fwd_err(got.cpu().numpy(), ref) <= FWD_TOL


def test_tile_agrees_with_fwd5_at_the_headline_config(skb):
    """BASELINE configs[2] (128 x 128, len 64, dim 5, dyadic 2): the tile kernel agrees with fwd5_kernel (the default
    plan) to 1e-12 and with the reference fixture (leading 8 x 8 block) to the north_star tolerance."""
    g = torch.Generator().manual_seed(0)
    X = torch.rand((128, 64, 5), dtype=torch.float64, generator=g)
    Y = torch.rand((128, 64, 5), dtype=torch.float64, generator=g)
    lib = skb._lib.lib
    G5 = skb.ops.sigkernel_forward(X.cuda(), Y.cuda(), "rbf", 0.5, 2, "gram")
    lib.skb_set_tile_mode(1)
    try:
        Gt = skb.ops.sigkernel_forward(X.cuda(), Y.cuda(), "rbf", 0.5, 2, "gram")
        Gs = skb.ops.sigkernel_forward(X[32:64].cuda(), Y[5:7].cuda(), "rbf", 0.5, 2, "gram")
    finally:
        lib.skb_set_tile_mode(-1)
    assert fwd_err(Gt.cpu().numpy(), G5.cpu().numpy()) <= 1e-12
    meta, z = load_golden("cfg3_gram_rbf")
    assert np.array_equal(X[:8].numpy(), z["X"])
    assert fwd_err(Gt[:8, :8].cpu().numpy(), z["G"]) <= FWD_TOL
    # row-block invariance: a tile computed inside a bigger batch equals the same pairs alone
    assert torch.equal(Gs, Gt[32:64, 5:7])


def test_tile_overflow_stays_in_its_pair(skb, tile_on):
    """A pair whose PDE solution overflows returns inf / nan and leaves the other 31 lanes of its tile alone."""
    X = make_paths("rand", 5, (34, 12, 2))
    Y = make_paths("rand", 6, (2, 12, 2))
    Xb = X.clone()
    Xb[3] *= 1e6
    Xb[3, ::2] *= -1
    G = skb.ops.sigkernel_forward(X.cuda(), Y.cuda(), "linear", 1.0, 1, "gram")
    Gb = skb.ops.sigkernel_forward(Xb.cuda(), Y.cuda(), "linear", 1.0, 1, "gram")
    keep = [i for i in range(34) if i != 3]
    assert torch.equal(G[keep], Gb[keep])
    assert not torch.isfinite(Gb[3]).all()
