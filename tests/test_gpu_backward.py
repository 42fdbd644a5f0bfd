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
                         make_paths)

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
    assert grad_err(Xd.grad.cpu().numpy(), expect.cpu().numpy()) <= 1e-12
