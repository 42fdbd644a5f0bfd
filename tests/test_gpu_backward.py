"""GPU parity tests of the adjoint (backward) path (run on the B200 box: `pytest -m gpu`).

The CUDA backward uses the analytic static-kernel derivative; the reference uses a one-sided finite
difference with h = 1e-9 whose own noise is ~1e-6 relative.  Hence two gates:
  * vs the reference's golden gradients:            max-norm relative error <= 5e-6
  * vs the oracle's analytic restatement (oracle #2): <= 1e-9
"""
import numpy as np
import pytest
import torch

from tests._util import (FWD_TOL, GRAD_TOL_ANALYTIC, GRAD_TOL_REF, fwd_err, golden_names, grad_err, load_golden,
                         make_paths, static_of)

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


@pytest.mark.parametrize("name", golden_names(("gram_bwd", "gram_sym_bwd", "kernel_bwd", "mmd_bwd")))
def test_backward_matches_reference_golden(skb, name):
    meta, z = load_golden(name)
    X, Y = _dev(z["X"]).requires_grad_(True), _dev(z["Y"])
    sk = skb.SigKernel(_static(skb, meta), meta["dyadic_order"], _naive_solver=meta["naive"])
    op = meta["op"]
    if op == "kernel_bwd":
        K = sk.compute_kernel(X, Y)
        assert fwd_err(K.detach().cpu().numpy(), z["K"]) <= FWD_TOL
        (K * _dev(z["w"])).sum().backward()
    elif op in ("gram_bwd", "gram_sym_bwd"):
        sym = op == "gram_sym_bwd"
        G = sk.compute_Gram(X, X if sym else Y, sym=sym)
        assert fwd_err(G.detach().cpu().numpy(), z["G"]) <= FWD_TOL
        (G * _dev(z["w"])).sum().backward()
    else:
        m = sk.compute_mmd(X, Y)
        assert fwd_err(m.detach().cpu().numpy(), z["mmd"]) <= FWD_TOL
        m.backward()
    assert X.grad.shape == X.shape
    assert grad_err(X.grad.cpu().numpy(), z["grad"]) <= GRAD_TOL_REF


SHAPES = [
    # A, B, M, N, D, d, kind
    (2, 3, 2, 2, 1, 0, "rand"),
    (2, 2, 2, 5, 2, 2, "rand"),
    (3, 2, 7, 6, 3, 0, "rand"),
    (3, 4, 9, 7, 3, 1, "rand"),
    (2, 2, 33, 17, 2, 1, "bm"),
    (2, 3, 64, 64, 3, 1, "rand"),
    (1, 2, 64, 40, 5, 2, "rand"),
    (2, 1, 40, 70, 8, 0, "bm"),
    (1, 2, 100, 20, 4, 1, "bm"),
    (2, 2, 12, 12, 11, 3, "bm"),
]


@pytest.mark.parametrize("A,B,M,N,D,d,kind", SHAPES)
@pytest.mark.parametrize("static", ["rbf", "linear"])
def test_grad_points_vs_analytic_oracle(skb, O, A, B, M, N, D, d, kind, static):
    X = make_paths(kind, 300 + M, (A, M, D))
    Y = make_paths(kind, 400 + N, (B, N, D))
    ok = O.RBFKernel(0.7) if static == "rbf" else O.LinearKernel()
    Gref, gp_ref, _ = O.gram_grad_points_analytic(X, Y, ok, d)
    G, gp = skb.ops.sigkernel_forward_backward(X.cuda(), Y.cuda(), static, 0.7 if static == "rbf" else 1.0, d, "gram")
    assert gp.shape == (A, B, M, D)
    assert fwd_err(G.cpu().numpy(), Gref.numpy()) <= FWD_TOL
    assert grad_err(gp.cpu().numpy(), gp_ref.numpy()) <= GRAD_TOL_ANALYTIC


@pytest.mark.parametrize("naive", [False, True])
def test_batch_grad_points_vs_analytic_oracle(skb, O, naive):
    X, Y = make_paths("bm", 41, (5, 14, 3)), make_paths("bm", 42, (5, 9, 3))
    for static, ok, par in (("rbf", O.RBFKernel(1.3), 1.3), ("linear", O.LinearKernel(0.5), 0.25)):
        kref, gp_ref, _ = O.batch_grad_points_analytic(X, Y, ok, 2, naive)
        k, gp = skb.ops.sigkernel_forward_backward(X.cuda(), Y.cuda(), static, par, 2, "batch", naive)
        assert gp.shape == (5, 14, 3)
        assert fwd_err(k.cpu().numpy(), kref.numpy()) <= FWD_TOL
        assert grad_err(gp.cpu().numpy(), gp_ref.numpy()) <= GRAD_TOL_ANALYTIC


@pytest.mark.parametrize("d", [0, 1, 2])
def test_sensitivity_from_static_vs_oracle(skb, O, d):
    X, Y = make_paths("rand", 51, (3, 11, 2)), make_paths("rand", 52, (2, 8, 2))
    sk = O.RBFKernel(0.5)
    Gref, _, Sref = O.gram_grad_points_analytic(X, Y, sk, d)
    G, S = skb.ops.sensitivity_from_static(sk.Gram_matrix(X, Y).cuda(), d, "gram")
    assert S.shape == (3, 2, 10, 7)
    assert fwd_err(G.cpu().numpy(), Gref.numpy()) <= FWD_TOL
    assert grad_err(S.cpu().numpy(), Sref.numpy()) <= 1e-11


def test_sensitivity_from_static_of_a_long_path_vs_oracle(skb, O):
    """Plugin path beyond the register-resident kernels (materialised grids): forward value and coarse sensitivities."""
    X, Y = make_paths("bm", 53, (2, 400, 2)), make_paths("bm", 54, (2, 9, 2))
    sk = O.RBFKernel(0.9)
    for d, naive in ((1, False), (2, True)):
        Gref, _, Sref = O.gram_grad_points_analytic(X, Y, sk, d, naive=naive)
        G, S = skb.ops.sensitivity_from_static(sk.Gram_matrix(X, Y).cuda(), d, "gram", naive)
        assert S.shape == (2, 2, 399, 8)
        assert fwd_err(G.cpu().numpy(), Gref.numpy()) <= FWD_TOL
        assert grad_err(S.cpu().numpy(), Sref.numpy()) <= 1e-11


def test_plugin_kernel_gradient_matches_reference_formula(skb, O):
    """User-defined static kernel: finite-difference route (as the reference) on top of the CUDA S."""
    class Poly:
        def batch_kernel(self, X, Y):
            return (1. + torch.bmm(X, Y.transpose(1, 2))) ** 2

        def Gram_matrix(self, X, Y):
            return (1. + torch.einsum('ipk,jqk->ijpq', X, Y)) ** 2

    X, Y = make_paths("bm", 61, (3, 10, 2)), make_paths("bm", 62, (4, 8, 2))
    w = torch.linspace(0.5, 1.5, 12, dtype=torch.float64).reshape(3, 4)
    _, gp_ref = O.gram_grad_points(X, Y, Poly(), 1)
    gref = O.gram_vjp(w, gp_ref)
    Xd = X.cuda().requires_grad_(True)
    G = skb.SigKernel(Poly(), 1).compute_Gram(Xd, Y.cuda())
    (G * w.cuda()).sum().backward()
    assert grad_err(Xd.grad.cpu().numpy(), gref.numpy()) <= GRAD_TOL_REF
    # batch form
    _, gpb_ref = O.batch_grad_points(X, Y[:3], Poly(), 0)
    Xb = X.cuda().requires_grad_(True)
    skb.SigKernel(Poly(), 0).compute_kernel(Xb, Y[:3].cuda()).sum().backward()
    assert grad_err(Xb.grad.cpu().numpy(), gpb_ref.numpy()) <= GRAD_TOL_REF


def test_chunked_workspace_gives_same_result(skb):
    """A workspace with room for only a few pairs' grids forces the chunked path: same bits."""
    lib, chk = skb._lib.lib, skb._lib.check
    X, Y = make_paths("rand", 71, (5, 9, 3)).cuda(), make_paths("rand", 72, (4, 7, 3)).cuda()
    A, M, D = X.shape
    B, N, _ = Y.shape
    G0, gp0 = skb.ops.sigkernel_forward_backward(X, Y, "rbf", 0.5, 1, "gram")
    full = lib.skb_bwd_workspace_bytes(A, B, M, N, D, 1, 0)
    one = lib.skb_bwd_workspace_bytes(1, 1, M, N, D, 1, 0)
    small = full - (A * B - 3) * ((N - 1) * 2 * 32 * 2 * 8)      # room for ~3 grids
    assert small > 0 and one > 0
    out = torch.empty(A * B, dtype=torch.float64, device="cuda")
    gp = torch.empty((A * B, M, D), dtype=torch.float64, device="cuda")
    ws = torch.empty(small, dtype=torch.uint8, device="cuda")
    chk(lib.skb_sigkernel_fwd_bwd(X.data_ptr(), Y.data_ptr(), 0, A, B, M, N, D, 1, 1, 0.5, 0, 0,
                                  out.data_ptr(), gp.data_ptr(), ws.data_ptr(), small,
                                  torch.cuda.current_stream().cuda_stream))
    assert torch.equal(out.view(A, B), G0) and torch.equal(gp.view(A, B, M, D), gp0)


def test_cfg4_shape_properties(skb, O):
    """compute_mmd + backward at the BASELINE cfg4 shape (128 x 64 x 3, dyadic 1, RBF): rows of
    grad_points are independent of the batch they are computed in; a sub-block matches the oracle."""
    g = torch.Generator().manual_seed(0)
    X = torch.rand((128, 64, 3), dtype=torch.float64, generator=g)
    Y = torch.rand((128, 64, 3), dtype=torch.float64, generator=g)
    G, gp = skb.ops.sigkernel_forward_backward(X.cuda(), Y.cuda(), "rbf", 0.5, 1, "gram")
    Gs, gps = skb.ops.sigkernel_forward_backward(X[5:8].cuda(), Y[100:102].cuda(), "rbf", 0.5, 1, "gram")
    assert torch.equal(G[5:8, 100:102], Gs) and torch.equal(gp[5:8, 100:102], gps)
    Gref, gp_ref, _ = O.gram_grad_points_analytic(X[5:8], Y[100:102], O.RBFKernel(0.5), 1)
    assert fwd_err(Gs.cpu().numpy(), Gref.numpy()) <= FWD_TOL
    assert grad_err(gps.cpu().numpy(), gp_ref.numpy()) <= GRAD_TOL_ANALYTIC
    Xd = X.cuda().requires_grad_(True)
    m = skb.SigKernel(skb.RBFKernel(0.5), 1).compute_mmd(Xd, Y.cuda())
    m.backward()
    assert torch.isfinite(Xd.grad).all() and Xd.grad.shape == (128, 64, 3)
    sk = skb.SigKernel(skb.RBFKernel(0.5), 1)
    Yd = Y.cuda()
    Kxx, Kyy, Kxy = sk.compute_Gram(Xd.detach(), Xd.detach(), sym=True), sk.compute_Gram(Yd, Yd, sym=True), G
    n = 128
    ref = ((Kxx.sum() - Kxx.diag().sum()) + (Kyy.sum() - Kyy.diag().sum())) / (n * (n - 1.)) - 2. * Kxy.mean()
    assert abs(float(m.detach()) - float(ref)) <= 1e-12
    # the gradient of the MMD is the reference's contraction of grad_points with d mmd / d K
    _, gxx = skb.ops.sigkernel_forward_backward(Xd.detach(), Xd.detach(), "rbf", 0.5, 1, "gram")
    wxx = (torch.ones(n, n, dtype=torch.float64, device="cuda") - torch.eye(n, dtype=torch.float64, device="cuda")) / (n * (n - 1.))
    expect = 2 * torch.einsum('ab,abmd->amd', wxx, gxx) + torch.einsum('ab,abmd->amd', torch.full_like(wxx, -2. / (n * n)), gp)
    # (the fused loss head rebuilds the forward grids -- off-diagonal pairs to ~1e-10 of their size, the diagonal of K_XX
    #  carries no weight -- where the eager call above falls back to the stored grid because of that diagonal)
    assert grad_err(Xd.grad.cpu().numpy(), expect.cpu().numpy()) <= 1e-9


# ---- adjoint by reconstruction (MODE_FWD_EMIT + MODE_REV_RECON), lazy fused backward, loss heads ------------------------
@pytest.fixture()
def stored_grid_only(skb):
    skb._lib.lib.skb_set_adjoint_mode(0)
    yield
    skb._lib.lib.skb_set_adjoint_mode(-1)


LONG = [
    # A, B, M, N, D, d          two and four warps per pair, 16 and 32 lanes per pair, every dyadic order
    (2, 2, 300, 12, 2, 1), (1, 2, 600, 9, 3, 0), (1, 1, 1000, 7, 2, 0), (2, 1, 130, 20, 8, 1), (1, 2, 256, 5, 5, 2),
    (2, 2, 17, 6, 2, 3), (2, 2, 33, 40, 3, 2), (3, 2, 64, 64, 3, 1), (2, 2, 20, 9, 11 - 2, 0),
]


@pytest.mark.parametrize("A,B,M,N,D,d", LONG)
@pytest.mark.parametrize("static", ["rbf", "linear"])
def test_reconstruction_adjoint_vs_analytic_oracle(skb, O, A, B, M, N, D, d, static):
    X = make_paths("bm", 500 + M, (A, M, D))
    Y = make_paths("bm", 600 + N, (B, N, D))
    ok = O.RBFKernel(1.1) if static == "rbf" else O.LinearKernel()
    par = 1.1 if static == "rbf" else 1.0
    assert skb.ops.adjoint_plan(M, N, D, d, static) == 6
    Gref, gp_ref, _ = O.gram_grad_points_analytic(X, Y, ok, d)
    G, gp = skb.ops.sigkernel_forward_backward(X.cuda(), Y.cuda(), static, par, d, "gram")
    assert fwd_err(G.cpu().numpy(), Gref.numpy()) <= FWD_TOL
    assert grad_err(gp.cpu().numpy(), gp_ref.numpy()) <= GRAD_TOL_ANALYTIC
    # the lazy pair of entry points: same numbers, plus the fused contraction with an upstream gradient
    res = skb.ops.sigkernel_forward_ctx(X.cuda(), Y.cuda(), static, par, d, "gram")
    assert res is not None
    G2, bctx = res
    assert torch.equal(G2, G)
    w = torch.linspace(-1.0, 2.0, A * B, dtype=torch.float64).reshape(A, B).cuda()
    gx, gp2 = skb.ops.sigkernel_backward_vjp(X.cuda(), Y.cuda(), static, par, d, "gram", bctx, "gram", grad_out=w, want_points=True)
    assert torch.equal(gp2, gp)
    expect = torch.einsum('ab,abmd->amd', w, gp)
    assert grad_err(gx.cpu().numpy(), expect.cpu().numpy()) <= 1e-12


ANY_LENGTH = [
    # A, B, M, N, D, d, naive    beyond every register-resident adjoint kernel: materialised grids (skb_generic_adj.cu)
    (2, 2, 300, 12, 2, 2, False), (1, 2, 600, 9, 3, 1, False), (1, 1, 1000, 7, 2, 1, False), (2, 1, 1100, 5, 4, 0, False),
    (1, 2, 70, 11, 3, 4, False), (2, 2, 300, 6, 2, 2, True),
]


@pytest.mark.parametrize("A,B,M,N,D,d,naive", ANY_LENGTH)
@pytest.mark.parametrize("static", ["rbf", "linear"])
def test_backward_of_any_length_vs_analytic_oracle(skb, O, A, B, M, N, D, d, naive, static):
    X = make_paths("bm", 700 + M, (A, M, D))
    Y = make_paths("bm", 800 + N, (B, N, D))
    ok = O.RBFKernel(0.9) if static == "rbf" else O.LinearKernel()
    par = 0.9 if static == "rbf" else 1.0
    assert skb.ops.adjoint_plan(M, N, D, d, static, naive) == 7
    Gref, gp_ref, _ = O.gram_grad_points_analytic(X, Y, ok, d, naive=naive)
    G, gp = skb.ops.sigkernel_forward_backward(X.cuda(), Y.cuda(), static, par, d, "gram", naive)
    assert fwd_err(G.cpu().numpy(), Gref.numpy()) <= FWD_TOL
    assert grad_err(gp.cpu().numpy(), gp_ref.numpy()) <= GRAD_TOL_ANALYTIC
    # and through autograd: d (sum w G) / d X
    Xg = X.cuda().requires_grad_(True)
    sk = skb.SigKernel(skb.RBFKernel(0.9) if static == "rbf" else skb.LinearKernel(), d, _naive_solver=naive)
    w = torch.linspace(-1.0, 2.0, A * B, dtype=torch.float64).reshape(A, B)
    (sk.compute_Gram(Xg, Y.cuda()) * w.cuda()).sum().backward()
    expect = torch.einsum('ab,abmd->amd', w, gp_ref)
    assert grad_err(Xg.grad.cpu().numpy(), expect.numpy()) <= GRAD_TOL_ANALYTIC


def test_batch_backward_of_any_length_vs_analytic_oracle(skb, O):
    """compute_kernel (pairs = batch) beyond the register-resident adjoint kernels: the materialised-grid path."""
    X, Y = make_paths("bm", 71, (3, 700, 2)), make_paths("bm", 72, (3, 8, 2))
    for static, ok, par in (("rbf", O.RBFKernel(1.3), 1.3), ("linear", O.LinearKernel(0.5), 0.25)):
        assert skb.ops.adjoint_plan(700, 8, 2, 1, static) == 7
        kref, gp_ref, _ = O.batch_grad_points_analytic(X, Y, ok, 1)
        k, gp = skb.ops.sigkernel_forward_backward(X.cuda(), Y.cuda(), static, par, 1, "batch")
        assert gp.shape == (3, 700, 2)
        assert fwd_err(k.cpu().numpy(), kref.numpy()) <= FWD_TOL
        assert grad_err(gp.cpu().numpy(), gp_ref.numpy()) <= GRAD_TOL_ANALYTIC
    Xg = X.cuda().requires_grad_(True)
    skb.SigKernel(skb.RBFKernel(1.3), 1).compute_kernel(Xg, Y.cuda()).sum().backward()
    kref, gp_ref, _ = O.batch_grad_points_analytic(X, Y, O.RBFKernel(1.3), 1)
    assert grad_err(Xg.grad.cpu().numpy(), gp_ref.numpy()) <= GRAD_TOL_ANALYTIC


@pytest.mark.parametrize("A,M,D,d,static", [(9, 20, 3, 1, "rbf"), (6, 64, 3, 1, "rbf"), (5, 33, 5, 2, "rbf"), (4, 60, 2, 0, "linear"),
                                            (7, 12, 2, 3, "rbf"), (3, 64, 2, 0, "rbf"), (4, 30, 8, 1, "linear")])
def test_unordered_pair_sweep_matches_the_full_square(skb, O, A, M, D, d, static):
    """Gram(X, X) in a loss head: one reversed sweep per unordered pair (d k / d X_a and d k / d X_b from the same
    sensitivities) against the two-sweeps-per-pair path (adjoint mode 3) and against the oracle's analytic gradient."""
    X = make_paths("bm", 900 + M, (A, M, D))
    par = 0.8 if static == "rbf" else 1.0
    assert skb.ops.adjoint_sym_supported(M, D, d, static)
    assert not skb.ops.adjoint_sym_supported(100, 2, 0, static)      # strips whose sums do not fit registers: two sweeps per pair
    # (loss heads weigh the diagonal of Gram(X, X) with zero; the diagonal pairs k(X_a, X_a) are the fastest-growing grids, where the
    #  reconstruction is accurate to ~1e-8 rather than 1e-12: checked separately below with the tolerance that goes with it)
    w_diag, w_off = 0.0, -0.7
    Xc = X.cuda()
    res = skb.ops.sigkernel_forward_ctx(Xc, Xc, static, par, d, "sym")
    assert res is not None
    G, bctx = res
    g_sym = skb.ops.sigkernel_backward_vjp(Xc, Xc, static, par, d, "sym", bctx, "sym", w_diag=w_diag, w_off=w_off, out_scale=2.0)
    # oracle: out_scale * sum_b coef(a,b) d1 k(X_a, X_b)
    ok = O.RBFKernel(par) if static == "rbf" else O.LinearKernel()
    _, gp_ref, _ = O.gram_grad_points_analytic(X, X, ok, d)
    coef = torch.full((A, A), w_off, dtype=torch.float64) + (w_diag - w_off) * torch.eye(A, dtype=torch.float64)
    expect = 2.0 * torch.einsum('ab,abmd->amd', coef, gp_ref)
    assert grad_err(g_sym.cpu().numpy(), expect.numpy()) <= GRAD_TOL_ANALYTIC
    g_diag = skb.ops.sigkernel_backward_vjp(Xc, Xc, static, par, d, "sym", bctx, "sym", w_diag=1.0, w_off=0.0)
    expect = torch.einsum('aamd->amd', gp_ref)
    assert grad_err(g_diag.cpu().numpy(), expect.numpy()) <= 1e-6
    # a general (asymmetric) upstream gradient: g[a] = sum_b w[a,b] d1 k(X_a, X_b), the term of (b,a) from the sweep of (a,b)
    w = torch.linspace(-1.0, 2.0, A * A, dtype=torch.float64).reshape(A, A)
    w = w - torch.diag(torch.diag(w))            # (the diagonal pairs are checked above with their own tolerance)
    g_w = skb.ops.sigkernel_backward_vjp(Xc, Xc, static, par, d, "sym", bctx, "sym", grad_out=w.cuda())
    expect = torch.einsum('ab,abmd->amd', w, gp_ref)
    assert grad_err(g_w.cpu().numpy(), expect.numpy()) <= GRAD_TOL_ANALYTIC
    # the eager entry point with pairs = sym: the full (A, A) / (A, A, M, D) tensors from the triangle of sweeps
    G2, gp2 = skb.ops.sigkernel_forward_backward(Xc, Xc, static, par, d, "sym")
    Gg, gpg = skb.ops.sigkernel_forward_backward(Xc, Xc, static, par, d, "gram")
    assert fwd_err(G2.cpu().numpy(), Gg.cpu().numpy()) <= 1e-13
    off = ~torch.eye(A, dtype=torch.bool)
    assert grad_err(gp2.cpu()[off].numpy(), gp_ref[off].numpy()) <= GRAD_TOL_ANALYTIC
    assert grad_err(gp2.cpu().numpy(), gpg.cpu().numpy()) <= 1e-6           # (diagonal pairs: see above)
    # and the public loss head with the sweep switched off (mode 3) gives the same gradient
    Y = make_paths("bm", 901 + M, (A + 1, M, D)).cuda()
    sk = skb.SigKernel(skb.RBFKernel(par) if static == "rbf" else skb.LinearKernel(), d)
    grads = []
    for mode in (-1, 3):
        skb._lib.lib.skb_set_adjoint_mode(mode)
        try:
            Xg = Xc.clone().requires_grad_(True)
            sk.compute_mmd(Xg, Y).backward()
            grads.append(Xg.grad.clone())
        finally:
            skb._lib.lib.skb_set_adjoint_mode(-1)
    assert grad_err(grads[0].cpu().numpy(), grads[1].cpu().numpy()) <= 1e-9


def test_reconstruction_agrees_with_the_stored_grid_kernels(skb):
    X, Y = make_paths("rand", 81, (6, 40, 3)).cuda(), make_paths("rand", 82, (5, 33, 3)).cuda()
    lib = skb._lib.lib
    G1, gp1 = skb.ops.sigkernel_forward_backward(X, Y, "rbf", 0.5, 1, "gram")
    lib.skb_set_adjoint_mode(0)
    try:
        G0, gp0 = skb.ops.sigkernel_forward_backward(X, Y, "rbf", 0.5, 1, "gram")
    finally:
        lib.skb_set_adjoint_mode(-1)
    lib.skb_set_adjoint_mode(2)          # 16 lanes per pair
    try:
        G2, gp2 = skb.ops.sigkernel_forward_backward(X, Y, "rbf", 0.5, 1, "gram")
    finally:
        lib.skb_set_adjoint_mode(-1)
    assert fwd_err(G1.cpu().numpy(), G0.cpu().numpy()) <= 1e-12 and fwd_err(G2.cpu().numpy(), G0.cpu().numpy()) <= 1e-12
    assert grad_err(gp1.cpu().numpy(), gp0.cpu().numpy()) <= 1e-10
    assert grad_err(gp2.cpu().numpy(), gp0.cpu().numpy()) <= 1e-10


def test_unstable_reconstruction_falls_back_to_the_stored_grid(skb, O):
    """Increments large enough that the PDE solution grows by many orders of magnitude: rebuilding it backwards loses
    all accuracy (growth squared), the boundary check raises the flag and the stored-grid kernels queued behind it take
    over -- the result still matches the oracle; without the fallback room the flag is left for the caller."""
    X = make_paths("rand", 91, (3, 40, 3)) * 3.0
    Y = make_paths("rand", 92, (2, 40, 3)) * 3.0
    Gref, gp_ref, _ = O.gram_grad_points_analytic(X, Y, O.LinearKernel(), 1)
    assert float(Gref.abs().max()) > 1e8
    G, gp = skb.ops.sigkernel_forward_backward(X.cuda(), Y.cuda(), "linear", 1.0, 1, "gram")
    assert grad_err(G.cpu().numpy(), Gref.numpy()) <= 1e-10
    assert grad_err(gp.cpu().numpy(), gp_ref.numpy()) <= GRAD_TOL_ANALYTIC
    # fused loss head through the same fallback
    res = skb.ops.sigkernel_forward_ctx(X.cuda(), Y.cuda(), "linear", 1.0, 1, "gram")
    w = torch.linspace(0.5, 1.5, 6, dtype=torch.float64).reshape(3, 2)
    gx = skb.ops.sigkernel_backward_vjp(X.cuda(), Y.cuda(), "linear", 1.0, 1, "gram", res[1], "gram", grad_out=w.cuda())
    assert grad_err(gx.cpu().numpy(), O.gram_vjp(w, gp_ref).numpy()) <= GRAD_TOL_ANALYTIC
    # a workspace without room for a forward grid: no fallback is queued, the flag word (byte 64) is set
    lib, chk = skb._lib.lib, skb._lib.check
    Xc, Yc = X.cuda(), Y.cuda()
    A, M, D = Xc.shape
    B, N, _ = Yc.shape
    Dp = 4
    fixed = 256 + 2 * (((A * M * Dp * 8) + 255) // 256 * 256) + 2 * (((B * N * Dp * 8) + 255) // 256 * 256) + lib.skb_ctx_bytes(A, B, M, N, 1, 0)
    ws = torch.zeros(fixed + 512, dtype=torch.uint8, device="cuda")
    out = torch.empty(A * B, dtype=torch.float64, device="cuda")
    gpo = torch.empty((A * B, M, D), dtype=torch.float64, device="cuda")
    chk(lib.skb_sigkernel_fwd_bwd(Xc.data_ptr(), Yc.data_ptr(), 0, A, B, M, N, D, 1, 0, 1.0, 0, 0, out.data_ptr(), gpo.data_ptr(),
                                  ws.data_ptr(), fixed + 512, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert int(ws[64:68].view(torch.int32).item()) == 1
    # and a benign problem leaves it clear
    Xs, Ys = (X / 3.0).cuda(), (Y / 3.0).cuda()
    chk(lib.skb_sigkernel_fwd_bwd(Xs.data_ptr(), Ys.data_ptr(), 0, A, B, M, N, D, 1, 0, 1.0, 0, 0, out.data_ptr(), gpo.data_ptr(),
                                  ws.data_ptr(), fixed + 512, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert int(ws[64:68].view(torch.int32).item()) == 0


def test_symmetric_gram_with_gradients(skb, O):
    """compute_Gram(X, X, sym=True) with X.requires_grad: same values and gradient as sym=False; the reference doubles
    the gradient because Y (= X) requires grad (sigkernel.py:410-412).  The ABI can also run the reversed sweep from the
    boundaries of a triangular forward (the pair (a, b), a > b, reads the transposed grid of (b, a))."""
    X = make_paths("rand", 95, (7, 21, 3))
    w = torch.rand(7, 7, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    w = 0.5 * (w + w.T)
    sk = skb.SigKernel(skb.RBFKernel(0.5), 1)
    Xs = X.cuda().requires_grad_(True)
    Gs = sk.compute_Gram(Xs, Xs, sym=True)
    (Gs * w.cuda()).sum().backward()
    Xf = X.cuda().requires_grad_(True)
    Gf = sk.compute_Gram(Xf, Xf, sym=False)
    (Gf * w.cuda()).sum().backward()
    assert fwd_err(Gs.detach().cpu().numpy(), Gf.detach().cpu().numpy()) <= 1e-12
    assert grad_err(Xs.grad.cpu().numpy(), Xf.grad.cpu().numpy()) <= 1e-11
    _, gp_ref, _ = O.gram_grad_points_analytic(X, X, O.RBFKernel(0.5), 1)
    gref = O.gram_vjp(w, gp_ref, y_requires_grad=True)
    assert grad_err(Xs.grad.cpu().numpy(), gref.numpy()) <= GRAD_TOL_ANALYTIC
    # triangular forward + reversed sweep over the square, straight through ops
    Xc = X.cuda()
    Gt, bctx = skb.ops.sigkernel_forward_ctx(Xc, Xc, "rbf", 0.5, 1, "sym")
    assert fwd_err(Gt.cpu().numpy(), Gf.detach().cpu().numpy()) <= 1e-12
    gx = skb.ops.sigkernel_backward_vjp(Xc, Xc, "rbf", 0.5, 1, "gram", bctx, "sym", grad_out=w.cuda(), out_scale=2.0)
    assert grad_err(gx.cpu().numpy(), gref.numpy()) <= 1e-8
    # sym=True with two different tensors is NOT symmetric: it must not take the triangular shortcut
    Y = make_paths("rand", 96, (7, 21, 3)).cuda()
    G_bad = sk.compute_Gram(X.cuda(), Y, sym=True)
    assert fwd_err(G_bad.cpu().numpy(), sk.compute_Gram(X.cuda(), Y).cpu().numpy()) <= 1e-13


@pytest.mark.parametrize("which", ["mmd", "score", "distance"])
def test_fused_loss_heads_match_the_composition(skb, O, which):
    """compute_mmd / compute_scoring_rule / compute_distance through the fused loss head (_SigLoss) vs the reference's
    composition of Grams evaluated with the oracle's analytic gradients."""
    n, m = (6, 5) if which != "distance" else (6, 6)
    X, Y = make_paths("rand", 101, (n, 18, 3)), make_paths("rand", 102, (m, 15, 3))
    ok, d = O.RBFKernel(0.6), 1
    sk = skb.SigKernel(skb.RBFKernel(0.6), d)
    Xd = X.cuda().requires_grad_(True)
    fn = {"mmd": sk.compute_mmd, "score": sk.compute_expected_scoring_rule, "distance": sk.compute_distance}[which]
    loss = fn(Xd, Y.cuda())
    assert loss.dim() == 0 and loss.grad_fn is not None and "SigLoss" in type(loss.grad_fn).__name__
    (3.0 * loss).backward()
    if which == "distance":
        kxx, gxx, _ = O.batch_grad_points_analytic(X, X, ok, d)
        kyy = O.compute_kernel(Y, Y, ok, d)
        kxy, gxy, _ = O.batch_grad_points_analytic(X, Y, ok, d)
        ref = kxx.mean() + kyy.mean() - 2 * kxy.mean()
        gref = 3.0 * (gxx / n - 2.0 * gxy / n)
    else:
        Gxx, gxx, _ = O.gram_grad_points_analytic(X, X, ok, d)
        Gxy, gxy, _ = O.gram_grad_points_analytic(X, Y, ok, d)
        wxx = (torch.ones(n, n, dtype=torch.float64) - torch.eye(n, dtype=torch.float64)) / (n * (n - 1.))
        wxy = torch.full((n, m), -2. / (n * m), dtype=torch.float64)
        ref = O._offdiag_mean(Gxx) - 2. * Gxy.mean()
        if which == "mmd":
            ref = ref + O._offdiag_mean(O.compute_Gram(Y, Y, ok, d))
        gref = 3.0 * (O.gram_vjp(wxx, gxx, True) + O.gram_vjp(wxy, gxy, False))
    assert abs(float(loss.detach()) - float(ref)) <= 1e-11 * (abs(float(ref)) + 1)
    assert grad_err(Xd.grad.cpu().numpy(), gref.numpy()) <= GRAD_TOL_ANALYTIC
    # no-grad evaluation takes the same head and returns the same number
    with torch.no_grad():
        assert abs(float(fn(Xd, Y.cuda())) - float(loss.detach())) <= 1e-13


def test_grad_mode_and_y_only_requires_grad(skb):
    """ Only Y requires grad: the reference returns no gradient for Y, and nothing here may crash;
    torch.no_grad() computes no backward work and returns tensors that do not require grad."""
    sk = skb.SigKernel(skb.RBFKernel(0.5), 1)
    X = make_paths("rand", 111, (3, 9, 2)).cuda()
    Y = make_paths("rand", 112, (4, 7, 2)).cuda().requires_grad_(True)
    G = sk.compute_Gram(X, Y)
    G.sum().backward()
    assert Y.grad is None
    K = sk.compute_kernel(X, Y[:3])
    K.sum().backward()
    assert Y.grad is None
    Xg = X.clone().requires_grad_(True)
    with torch.no_grad():
        G0 = sk.compute_Gram(Xg, Y.detach())
    assert not G0.requires_grad
    assert fwd_err(G0.cpu().numpy(), G.detach().cpu().numpy()) <= 1e-13


def test_function_space_kernels_vs_oracle_and_gradient_flow(skb, O):
    """RBF_ID / Linear_ID / RBF_CEXP (reference static_kernels.py:75-206): 4-D paths go through a host-side transform and
    the fused kernels; values vs the oracle fed the same kernels, gradients flow back through the transform."""
    X = make_paths("rand", 121, (3, 9, 4, 2))
    Y = make_paths("rand", 122, (4, 7, 4, 2))
    for mk, mo in ((skb.RBF_ID_Kernel(2.0), O.RBF_ID_Kernel(2.0)), (skb.Linear_ID_Kernel(), O.Linear_ID_Kernel()),
                   (skb.RBF_CEXP_Kernel(1.0, 2.0, 4), O.RBF_CEXP_Kernel(1.0, 2.0, 4))):
        ref = O.compute_Gram(X, Y, mo, 1).numpy()
        got = skb.SigKernel(mk, 1).compute_Gram(X.cuda(), Y.cuda())
        assert fwd_err(got.cpu().numpy(), ref) <= FWD_TOL
        kref = O.compute_kernel(X, Y[:3], mo, 0).numpy()
        kgot = skb.SigKernel(mk, 0).compute_kernel(X.cuda(), Y[:3].cuda())
        assert fwd_err(kgot.cpu().numpy(), kref) <= FWD_TOL
        # gradient w.r.t. the 4-D input = gradient w.r.t. the transformed paths pulled back through the transform
        Xd = X.cuda().requires_grad_(True)
        skb.SigKernel(mk, 1).compute_Gram(Xd, Y.cuda()).sum().backward()
        Xt = mo.transform(X).detach().requires_grad_(True)
        _, gp_ref, _ = O.gram_grad_points_analytic(Xt.detach(), mo.transform(Y), O.RBFKernel(mo.sigma) if hasattr(mo, "sigma") else O.LinearKernel(), 1)
        gt = gp_ref.sum(dim=1)
        Xc = X.clone().requires_grad_(True)
        (mo.transform(Xc) * gt).sum().backward()
        assert Xd.grad.shape == X.shape
        assert grad_err(Xd.grad.cpu().numpy(), Xc.grad.numpy()) <= GRAD_TOL_ANALYTIC
