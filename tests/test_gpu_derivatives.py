"""GPU parity of the kernel + directional-derivative solver (SigKernel.compute_kernel_and_derivatives_Gram,
reference sigkernel.py:43-89, 504-593; cuda_backend.py:165-223) through the C ABI.

Tolerances: the reference forms the increments by finite differences in eps = 1e-4 of fp64 static kernels, so
inc_diff carries ~1e-12 and inc_diffdiff ~1e-8 of absolute rounding noise that depends on how exp() / the dot
products round.  torch evaluates the static kernels on the GPU here and on the CPU in the oracle (different exp
and einsum implementations), hence k to 1e-10, k_gamma to 1e-7 and k_gamma_gamma to 1e-3 (mixed abs/rel, the
metric of tests/_util.fwd_err); fed IDENTICAL static matrices the CUDA kernels reproduce the oracle bit for bit
(test_derivatives_bitwise_from_static)."""
import numpy as np
import pytest
import torch

from tests._util import FWD_TOL, fwd_err, golden_names, load_golden, make_paths

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
    return mod.RBFKernel(meta["param"]) if meta["static"] == "rbf" else mod.LinearKernel(meta["param"])


@pytest.mark.parametrize("name", golden_names(("deriv",)))
def test_derivatives_match_reference_golden(skb, name):
    meta, z = load_golden(name)
    X, Y, g = (torch.from_numpy(z[k]).cuda() for k in ("X", "Y", "gamma"))
    sk = skb.SigKernel(_static(skb, meta), meta["dyadic_order"])
    K, Kd, Kdd = sk.compute_kernel_and_derivatives_Gram(X, Y, g)
    assert fwd_err(K.cpu().numpy(), z["K"]) <= FWD_TOL
    assert fwd_err(Kd.cpu().numpy(), z["K_diff"]) <= 1e-7
    assert fwd_err(Kdd.cpu().numpy(), z["K_diffdiff"]) <= 1e-3


@pytest.mark.parametrize("A,B,M,N,D,d,static", [
    (3, 4, 9, 7, 3, 0, "rbf"), (2, 3, 17, 12, 2, 1, "rbf"), (4, 2, 33, 40, 4, 2, "linear"), (1, 1, 2, 2, 1, 3, "rbf"),
    (2, 2, 150, 97, 2, 2, "rbf"),
])
def test_derivatives_bitwise_from_static(skb, O, A, B, M, N, D, d, static):
    """Same static matrices in => the increment build and the three stencils are bit-identical to the oracle."""
    X = make_paths("bm", 40 + M, (A, M, D))
    Y = make_paths("bm", 41 + N, (B, N, D))
    gamma = make_paths("rand", 42, (A, M, D))
    eps = 1e-4
    sk = O.RBFKernel(0.7) if static == "rbf" else O.LinearKernel(1.0)
    K0, K1, K2 = sk.Gram_matrix(X, Y), sk.Gram_matrix(X + eps * gamma, Y), sk.Gram_matrix(X + 2. * eps * gamma, Y)
    ref = O.compute_kernel_and_derivatives_Gram(X, Y, gamma, sk, d, eps)
    skb._lib.lib.skb_set_deriv_mode(0)           # the diagonal kernel: the reference's operation order
    try:
        got = skb.ops.kernel_and_derivatives_from_static(K0.cuda(), K1.cuda(), K2.cuda(), d, eps)
    finally:
        skb._lib.lib.skb_set_deriv_mode(-1)
    for r, g in zip(ref, got):
        assert np.array_equal(g.cpu().numpy(), r.numpy())
    # default dispatch (the streaming kernel up to 256 fine rows): same algebra, sums factored and FMA-contracted
    got = skb.ops.kernel_and_derivatives_from_static(K0.cuda(), K1.cuda(), K2.cuda(), d, eps)
    for r, g in zip(ref, got):
        assert fwd_err(g.cpu().numpy(), r.numpy()) <= 1e-11


@pytest.mark.parametrize("A,B,M,N,D,d", [(5, 3, 64, 64, 3, 2), (2, 2, 33, 9, 2, 3), (3, 3, 257, 5, 2, 0), (2, 3, 129, 40, 4, 1),
                                         (2, 2, 6, 70, 2, 2), (1, 2, 40, 3, 2, 1)])
def test_streaming_derivative_kernel_vs_oracle(skb, O, A, B, M, N, D, d):
    """Every strip height (1, 2, 4, 8 rows per lane) and dyadic order of the streaming kernel against the oracle."""
    X = make_paths("bm", 60 + M, (A, M, D))
    Y = make_paths("bm", 61 + N, (B, N, D))
    gamma = make_paths("rand", 62, (A, M, D))
    eps = 1e-4
    sk = O.RBFKernel(0.9)
    K0, K1, K2 = sk.Gram_matrix(X, Y), sk.Gram_matrix(X + eps * gamma, Y), sk.Gram_matrix(X + 2. * eps * gamma, Y)
    ref = O.compute_kernel_and_derivatives_Gram(X, Y, gamma, sk, d, eps)
    got = skb.ops.kernel_and_derivatives_from_static(K0.cuda(), K1.cuda(), K2.cuda(), d, eps)
    for r, g in zip(ref, got):
        assert fwd_err(g.cpu().numpy(), r.numpy()) <= 1e-11


def test_derivative_is_the_eps_derivative_of_the_kernel(skb):
    """k_gamma ~ d/ds k(X + s gamma, Y) at s = 0: check against a central difference of compute_Gram."""
    X = make_paths("bm", 50, (3, 12, 2)).cuda()
    Y = make_paths("bm", 51, (4, 10, 2)).cuda()
    gamma = make_paths("rand", 52, (3, 12, 2)).cuda()
    sk = skb.SigKernel(skb.RBFKernel(1.0), 2)
    K, Kd, _ = sk.compute_kernel_and_derivatives_Gram(X, Y, gamma)
    h = 1e-5
    fd = (sk.compute_Gram(X + h * gamma, Y) - sk.compute_Gram(X - h * gamma, Y)) / (2 * h)
    assert fwd_err(K.cpu().numpy(), sk.compute_Gram(X, Y).cpu().numpy()) <= FWD_TOL
    assert fwd_err(Kd.cpu().numpy(), fd.cpu().numpy()) <= 5e-3     # eps = 1e-4 one-sided FD inside k_kgrad


def test_derivatives_errors(skb):
    K = torch.zeros((2, 2, 4, 4), dtype=torch.float64)
    with pytest.raises(skb.SigKernelB200Error):
        skb.ops.kernel_and_derivatives_from_static(K, K, K, 0, 1e-4)          # CPU tensors
    Kc = K.cuda()
    with pytest.raises(skb.SigKernelB200Error):
        skb.ops.kernel_and_derivatives_from_static(Kc, Kc[:, :1], Kc, 0, 1e-4)  # shape mismatch
