"""oracle/sigkernel_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's signature-kernel hot path (crispitagorico/sigkernel
@ 40a5831), used only as the parity checker by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under sigkernel_b200/ imports it.

What is restated, and from where (all paths relative to /root/reference):
  static kernels              sigkernel/static_kernels.py:11-73
  second difference + tile    sigkernel/sigkernel.py:217-218, 362-364, 607-613
  PDE solve (S2 / S1)         sigkernel/cython_backend.pyx:7-33, 64-119   (oracle/solver.c, or
                              the compiled reference in oracle/_ref when present)
  adjoint backward            sigkernel/sigkernel.py:256-343 (batch), 419-502 (Gram)
  autograd glue               sigkernel/sigkernel.py:405-416
  MMD / distance / scoring    sigkernel/sigkernel.py:130-197

Pinning: the reference has no golden vectors of its own, so parity is
pinned by (1) tests/golden/*.npz, generated in the build container by importing the
UNMODIFIED reference package (tests/golden/make_golden.py) and (2) a bit-for-bit
comparison of solver.c with the compiled reference solver (oracle/_ref) whenever that
library is present.  torch is used for einsum/bmm/exp because those are the reference's
own arithmetic for the static kernels.
"""
import ctypes
import glob
import importlib.util
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_H_FD = 1e-9  # finite-difference step of the reference backward, sigkernel.py:314,473


# ---------------------------------------------------------------------------------------
# solver back ends: plain-C restatement (always) and the compiled reference (if built)
# ---------------------------------------------------------------------------------------
_clib = None


def _c():
    global _clib
    if _clib is None:
        path = os.path.join(_HERE, "liboracle_solver.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle_solver.so missing: run `make -C oracle`")
        lib = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.skb_oracle_solve_batch.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, dp]
        lib.skb_oracle_solve_gram.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int, dp]
        lib.skb_oracle_solve_gram_corner.argtypes = [dp, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                                     ctypes.c_int, dp, dp]
        lib.skb_oracle_solve_derivatives.argtypes = [dp, dp, dp, ctypes.c_long, ctypes.c_int, ctypes.c_int, dp, dp]
        for f in (lib.skb_oracle_solve_batch, lib.skb_oracle_solve_gram,
                  lib.skb_oracle_solve_gram_corner, lib.skb_oracle_solve_derivatives):
            f.restype = None
        _clib = lib
    return _clib


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


_ref_mod = False


def ref_backend():
    """The reference's own compiled solver (oracle/_ref, built by `make -C oracle ref`), or None."""
    global _ref_mod
    if _ref_mod is False:
        cands = glob.glob(os.path.join(_HERE, "_ref", "cython_backend*.so"))
        if cands:
            spec = importlib.util.spec_from_file_location("cython_backend", cands[0])
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _ref_mod = mod
        else:
            _ref_mod = None
    return _ref_mod


def solve_batch(inc, naive=False, backend="c"):
    """inc (A,MM,NN) fp64 -> full grid (A,MM+1,NN+1).  cython_backend.pyx:7-33."""
    inc = np.ascontiguousarray(inc, dtype=np.float64)
    A, MM, NN = inc.shape
    if backend == "ref":
        return np.asarray(ref_backend().sigkernel_cython(inc, bool(naive)))
    if backend == "numpy":
        return np.stack([_solve_numpy(inc[a], naive) for a in range(A)]) if A else \
            np.zeros((0, MM + 1, NN + 1))
    K = np.zeros((A, MM + 1, NN + 1), dtype=np.float64)
    _c().skb_oracle_solve_batch(_dptr(inc), A, MM, NN, int(bool(naive)), _dptr(K))
    return K


def solve_gram(inc, sym=False, naive=False, backend="c"):
    """inc (A,B,MM,NN) fp64 -> full grid (A,B,MM+1,NN+1).  cython_backend.pyx:64-119."""
    inc = np.ascontiguousarray(inc, dtype=np.float64)
    A, B, MM, NN = inc.shape
    if backend == "ref":
        return np.asarray(ref_backend().sigkernel_Gram_cython(inc, bool(sym), bool(naive)))
    if backend == "numpy":
        K = np.zeros((A, B, MM + 1, NN + 1))
        for a in range(A):
            for b in range(B):
                K[a, b] = _solve_numpy(inc[a, b], naive)
        return K
    K = np.zeros((A, B, MM + 1, NN + 1), dtype=np.float64)
    _c().skb_oracle_solve_gram(_dptr(inc), A, B, MM, NN, int(bool(sym)), int(bool(naive)), _dptr(K))
    return K


def solve_gram_corner(inc, naive=False):
    """inc (P,MM,NN) -> u[MM,NN] per pair, two live rows only (bench-sized inputs)."""
    inc = np.ascontiguousarray(inc, dtype=np.float64)
    P, MM, NN = inc.shape
    out = np.empty(P, dtype=np.float64)
    scratch = np.empty(2 * (NN + 1), dtype=np.float64)
    _c().skb_oracle_solve_gram_corner(_dptr(inc), P, MM, NN, int(bool(naive)), _dptr(out),
                                      _dptr(scratch))
    return out


def _solve_numpy(g, naive):
    """Pure-Python loop of the same stencil (small cases only); cython_backend.pyx:110-117."""
    MM, NN = g.shape
    u = np.ones((MM + 1, NN + 1))
    for i in range(MM):
        for j in range(NN):
            x = g[i, j]
            if naive:
                u[i + 1, j + 1] = (u[i + 1, j] + u[i, j + 1]) * (1. + 0.5 * x) - u[i, j]
            else:
                u[i + 1, j + 1] = (u[i + 1, j] + u[i, j + 1]) * (1. + 0.5 * x + (1. / 12) * (x * x)) \
                    - u[i, j] * (1. - (1. / 12) * (x * x))
    return u


# ---------------------------------------------------------------------------------------
# static kernels (static_kernels.py:11-73) -- torch arithmetic, like the reference
# ---------------------------------------------------------------------------------------
class LinearKernel:
    """<x,y>; batch form scales both sides by `scale`, Gram form ignores it
    (static_kernels.py:24 vs :33 -- an inconsistency of the reference that is preserved)."""

    def __init__(self, scale=1.0):
        self.scale = scale

    def batch_kernel(self, X, Y):
        return torch.bmm(self.scale * X, self.scale * Y.permute(0, 2, 1))

    def Gram_matrix(self, X, Y):
        return torch.einsum('ipk,jqk->ijpq', X, Y)


class RBFKernel:
    """exp(-|x-y|^2 / sigma) evaluated as exp(-((-2 x.y) + (|x|^2 + |y|^2)) / sigma)
    (static_kernels.py:42-73: sigma, not 2 sigma^2; true division)."""

    def __init__(self, sigma):
        self.sigma = sigma

    def batch_kernel(self, X, Y):
        A, M, N = X.shape[0], X.shape[1], Y.shape[1]
        xs = torch.sum(X ** 2, dim=2)
        ys = torch.sum(Y ** 2, dim=2)
        dist = -2. * torch.bmm(X, Y.permute(0, 2, 1))
        dist += xs.reshape(A, M, 1) + ys.reshape(A, 1, N)
        return torch.exp(-dist / self.sigma)

    def Gram_matrix(self, X, Y):
        A, B, M, N = X.shape[0], Y.shape[0], X.shape[1], Y.shape[1]
        xs = torch.sum(X ** 2, dim=2)
        ys = torch.sum(Y ** 2, dim=2)
        dist = -2. * torch.einsum('ipk,jqk->ijpq', X, Y)
        dist += xs.reshape(A, 1, M, 1) + ys.reshape(1, B, 1, N)
        return torch.exp(-dist / self.sigma)


# function-space kernels (static_kernels.py:75-250): paths are (batch, len_t, len_x, dim); every one of them
# is a path transform in front of Linear / RBF
def cos_exp_kernel(x_y, n_freqs=5, sigma=1):
    """static_kernels.py:233-250."""
    cos_term = torch.cos(2 * torch.pi * x_y[:, :, None] * torch.arange(n_freqs)[None, None]).sum(dim=-1)
    return cos_term * torch.exp(-x_y ** 2 / sigma)


def CEXP(X, n_freqs=20, sigma=np.sqrt(10)):
    """static_kernels.py:208-231: integral operator of the cos-exp kernel along the len_x axis."""
    length_x = X.shape[2]
    grid = torch.linspace(0, 1, length_x, dtype=torch.float64)
    T_mat = cos_exp_kernel(grid[:, None] - grid[None, :], n_freqs=n_freqs, sigma=sigma)
    return ((1. / length_x) * torch.matmul(X.permute(0, 1, 3, 2), T_mat)).permute(0, 1, 3, 2)


def _flat(X):
    return X.reshape(X.shape[0], X.shape[1], -1)


class Linear_ID_Kernel(LinearKernel):
    """static_kernels.py:146-175."""

    def transform(self, X):
        return _flat(X)

    def batch_kernel(self, X, Y):
        return super().batch_kernel(_flat(X), _flat(Y))

    def Gram_matrix(self, X, Y):
        return super().Gram_matrix(_flat(X), _flat(Y))


class RBF_ID_Kernel(RBFKernel):
    """static_kernels.py:178-206."""

    def transform(self, X):
        return _flat(X)

    def batch_kernel(self, X, Y):
        return super().batch_kernel(_flat(X), _flat(Y))

    def Gram_matrix(self, X, Y):
        return super().Gram_matrix(_flat(X), _flat(Y))


class RBF_CEXP_Kernel(RBFKernel):
    """static_kernels.py:75-115."""

    def __init__(self, sigma1, sigma2, n_freqs):
        self.sigma1 = sigma1
        super().__init__(sigma2)
        self.n_freqs = n_freqs

    def transform(self, X):
        return _flat(CEXP(X, self.n_freqs, self.sigma1))

    def batch_kernel(self, X, Y):
        return super().batch_kernel(self.transform(X), self.transform(Y))

    def Gram_matrix(self, X, Y):
        return super().Gram_matrix(self.transform(X), self.transform(Y))


# ---------------------------------------------------------------------------------------
# increments: second difference + dyadic refinement (sigkernel.py:217-218, 362-364, 607-613)
# ---------------------------------------------------------------------------------------
def second_difference(K):
    """K (...,M,N) -> (...,M-1,N-1): ((K[i+1,j+1] + K[i,j]) - K[i+1,j]) - K[i,j+1]."""
    return K[..., 1:, 1:] + K[..., :-1, :-1] - K[..., 1:, :-1] - K[..., :-1, 1:]


def refine(inc_c, dyadic_order):
    """`tile(tile(.,r,2^d)/2^d, c, 2^d)/2^d`: each coarse cell becomes 2^d x 2^d fine cells,
    value divided by 2^d twice (exact power-of-two scaling)."""
    n = 2 ** dyadic_order
    t = torch.repeat_interleave(inc_c, n, dim=-2) / float(n)
    return torch.repeat_interleave(t, n, dim=-1) / float(n)


def increments(K_static, dyadic_order):
    return refine(second_difference(K_static), dyadic_order)


# ---------------------------------------------------------------------------------------
# forward: full solution grids
# ---------------------------------------------------------------------------------------
def batch_grid(X, Y, static_kernel, dyadic_order, naive=False, backend="c"):
    """sigkernel.py:204-253 (CPU branch).  Returns (grid (A,MM+1,NN+1), K_static (A,M,N))."""
    Ks = static_kernel.batch_kernel(X, Y)
    inc = increments(Ks, dyadic_order)
    return torch.from_numpy(solve_batch(inc.detach().numpy(), naive, backend)), Ks


def gram_grid(X, Y, static_kernel, dyadic_order, sym=False, naive=False, backend="c"):
    """sigkernel.py:349-401 (CPU branch).  Returns (grid (A,B,MM+1,NN+1), K_static (A,B,M,N))."""
    Ks = static_kernel.Gram_matrix(X, Y)
    inc = increments(Ks, dyadic_order)
    return torch.from_numpy(solve_gram(inc.detach().numpy(), sym, naive, backend)), Ks


def compute_kernel(X, Y, static_kernel, dyadic_order, naive=False, backend="c"):
    return batch_grid(X, Y, static_kernel, dyadic_order, naive, backend)[0][:, -1, -1]


def compute_Gram(X, Y, static_kernel, dyadic_order, sym=False, naive=False, backend="c"):
    return gram_grid(X, Y, static_kernel, dyadic_order, sym, naive, backend)[0][:, :, -1, -1]


def _offdiag_mean(K):
    n = K.shape[0]
    return (torch.sum(K) - torch.sum(torch.diag(K))) / (n * (n - 1.))


def compute_mmd(X, Y, static_kernel, dyadic_order, naive=False, backend="c"):
    """sigkernel.py:180-197."""
    Kxx = compute_Gram(X, X, static_kernel, dyadic_order, True, naive, backend)
    Kyy = compute_Gram(Y, Y, static_kernel, dyadic_order, True, naive, backend)
    Kxy = compute_Gram(X, Y, static_kernel, dyadic_order, False, naive, backend)
    return _offdiag_mean(Kxx) + _offdiag_mean(Kyy) - 2. * torch.mean(Kxy)


def compute_distance(X, Y, static_kernel, dyadic_order, naive=False, backend="c"):
    """sigkernel.py:130-144."""
    kxx = compute_kernel(X, X, static_kernel, dyadic_order, naive, backend)
    kyy = compute_kernel(Y, Y, static_kernel, dyadic_order, naive, backend)
    kxy = compute_kernel(X, Y, static_kernel, dyadic_order, naive, backend)
    return torch.mean(kxx) + torch.mean(kyy) - 2. * torch.mean(kxy)


def compute_scoring_rule(X, y, static_kernel, dyadic_order, naive=False, backend="c"):
    """sigkernel.py:146-178 (the expected variant is the same arithmetic)."""
    Kxx = compute_Gram(X, X, static_kernel, dyadic_order, True, naive, backend)
    Kxy = compute_Gram(X, y, static_kernel, dyadic_order, False, naive, backend)
    return _offdiag_mean(Kxx) - 2. * torch.mean(Kxy)


# ---------------------------------------------------------------------------------------
# backward, the reference way: reversed PDE + one-sided finite difference (h = 1e-9)
# ---------------------------------------------------------------------------------------
def _perturbed(X, D):
    """X (A,M,D) -> (A, M*D, D): row (p,c) is x_p + h e_c.  sigkernel.py:316-318, 475-477."""
    A, M = X.shape[0], X.shape[1]
    eye = torch.eye(D, dtype=X.dtype)
    Xh = X[:, :, None, :] + _H_FD * eye[None, None, :, :]      # (A,M,D(c),D)
    return Xh.reshape(A, M * D, D)


def _assemble_points(g1, g2):
    """g1,g2 (...,M-1,D) -> per-point gradient (...,M,D).  sigkernel.py:337-340, 497-500."""
    first = (g2[..., 0, :] - g1[..., 0, :])[..., None, :]
    mid = g1[..., :-1, :] + g2[..., 1:, :] - g1[..., 1:, :]
    last = g1[..., -1, :][..., None, :]
    return torch.cat([first, mid, last], dim=-2)


def gram_grad_points(X, Y, static_kernel, dyadic_order, sym=False, naive=False, backend="c"):
    """Reference `prep_backward` (sigkernel.py:419-502): returns (G (A,B), grad_points (A,B,M,D))."""
    A, B, M, N, D = X.shape[0], Y.shape[0], X.shape[1], Y.shape[1], X.shape[2]
    n = 2 ** dyadic_order
    U, Ks = gram_grid(X, Y, static_kernel, dyadic_order, sym, naive, backend)
    inc = increments(Ks, dyadic_order)
    inc_rev = torch.flip(inc, dims=[2, 3])
    U_rev = torch.from_numpy(solve_gram(inc_rev.numpy(), sym, naive, backend))
    U_rev = torch.flip(U_rev, dims=[2, 3])
    GG = U[:, :, :-1, :-1] * U_rev[:, :, 1:, 1:]                       # (A,B,MM,NN)

    Kh = static_kernel.Gram_matrix(_perturbed(X, D), Y)                # (A,B,M*D,N)
    Kh = Kh.reshape(A, B, M, D, N).permute(0, 1, 2, 4, 3)              # (A,B,M,N,D)
    d1 = Kh[:, :, 1:, 1:, :] - Kh[:, :, 1:, :-1, :] - Ks[:, :, 1:, 1:, None] + Ks[:, :, 1:, :-1, None]
    d2 = Kh[:, :, 1:, 1:, :] - Kh[:, :, 1:, :-1, :] - Ks[:, :, 1:, 1:, None] + Ks[:, :, 1:, :-1, None]
    d2 = d2 + (- Kh[:, :, :-1, 1:, :] + Kh[:, :, :-1, :-1, :] + Ks[:, :, :-1, 1:, None]
               - Ks[:, :, :-1, :-1, None])

    def tiled(d):
        t = torch.repeat_interleave(d, n, dim=2) / float(n)
        return torch.repeat_interleave(t, n, dim=3) / float(n)

    g1 = torch.sum((GG[..., None] * tiled(d1)) / _H_FD, dim=3)          # (A,B,MM,D)
    g1 = torch.sum(g1.reshape(A, B, M - 1, n, D), dim=3)
    g2 = torch.sum((GG[..., None] * tiled(d2)) / _H_FD, dim=3)
    g2 = torch.sum(g2.reshape(A, B, M - 1, n, D), dim=3)
    return U[:, :, -1, -1], _assemble_points(g1, g2)


def batch_grad_points(X, Y, static_kernel, dyadic_order, naive=False, backend="c"):
    """Reference `_SigKernel.backward` body (sigkernel.py:256-343): (k (A,), grad_points (A,M,D))."""
    A, M, N, D = X.shape[0], X.shape[1], Y.shape[1], X.shape[2]
    n = 2 ** dyadic_order
    U, Ks = batch_grid(X, Y, static_kernel, dyadic_order, naive, backend)
    inc = increments(Ks, dyadic_order)
    U_rev = torch.from_numpy(solve_batch(torch.flip(inc, dims=[1, 2]).numpy(), naive, backend))
    U_rev = torch.flip(U_rev, dims=[1, 2])
    KK = U[:, :-1, :-1] * U_rev[:, 1:, 1:]

    Kh = static_kernel.batch_kernel(_perturbed(X, D), Y)               # (A,M*D,N)
    Kh = Kh.reshape(A, M, D, N).permute(0, 1, 3, 2)                    # (A,M,N,D)
    d1 = Kh[:, 1:, 1:, :] - Kh[:, 1:, :-1, :] - Ks[:, 1:, 1:, None] + Ks[:, 1:, :-1, None]
    d2 = Kh[:, 1:, 1:, :] - Kh[:, 1:, :-1, :] - Ks[:, 1:, 1:, None] + Ks[:, 1:, :-1, None]
    d2 = d2 + (- Kh[:, :-1, 1:, :] + Kh[:, :-1, :-1, :] + Ks[:, :-1, 1:, None] - Ks[:, :-1, :-1, None])

    def tiled(d):
        t = torch.repeat_interleave(d, n, dim=1) / float(n)
        return torch.repeat_interleave(t, n, dim=2) / float(n)

    g1 = torch.sum((KK[..., None] * tiled(d1)) / _H_FD, dim=2)
    g1 = torch.sum(g1.reshape(A, M - 1, n, D), dim=2)
    g2 = torch.sum((KK[..., None] * tiled(d2)) / _H_FD, dim=2)
    g2 = torch.sum(g2.reshape(A, M - 1, n, D), dim=2)
    return U[:, -1, -1], _assemble_points(g1, g2)


def gram_vjp(grad_out, grad_points, y_requires_grad=False):
    """`_SigKernelGram.backward` (sigkernel.py:405-416): (A,B) x (A,B,M,D) -> (A,M,D),
    doubled when Y also requires grad (the reference assumes Y is X then)."""
    g = (grad_out[:, :, None, None] * grad_points).sum(dim=1)
    return 2 * g if y_requires_grad else g


# ---------------------------------------------------------------------------------------
# oracle #2: the same backward with the ANALYTIC static-kernel derivative (no h = 1e-9 noise)
# ---------------------------------------------------------------------------------------
def coarse_sensitivity(U, U_rev_flipped, dyadic_order):
    """S[i,j] = 4^-d * sum over the fine cells (p,q) of coarse cell (i,j) of u[p,q] u_rev[p+1,q+1]
    (closed form of A7).  U, U_rev_flipped (...,MM+1,NN+1) -> (...,M-1,N-1)."""
    n = 2 ** dyadic_order
    GG = U[..., :-1, :-1] * U_rev_flipped[..., 1:, 1:]
    sh = GG.shape
    Mc, Nc = sh[-2] // n, sh[-1] // n
    S = GG.reshape(*sh[:-2], Mc, n, Nc, n).sum(dim=(-3, -1))
    return S / float(n * n)


def _d1k(static_kernel, kind, x, y, Kxy):
    """d k(x,y) / d x for the two built-in kernels.  x (...,M,1,D), y (...,1,N,D), Kxy (...,M,N)."""
    if kind == "rbf":
        return (-2. / static_kernel.sigma) * (x - y) * Kxy[..., None]
    if kind == "linear_gram":
        return y.expand(*Kxy.shape, y.shape[-1])
    if kind == "linear_batch":
        return (static_kernel.scale ** 2) * y.expand(*Kxy.shape, y.shape[-1])
    raise ValueError(kind)


def grad_points_from_S(S, dk):
    """S (...,M-1,N-1), dk = d1k at all nodes (...,M,N,D) -> grad_points (...,M,D):
       hi[i,j] = dk[i+1,j+1] - dk[i+1,j];  lo[i,j] = dk[i,j] - dk[i,j+1];
       g[p] = sum_j S[p-1,j] hi[p-1,j] (p>=1) + S[p,j] lo[p,j] (p<=M-2)."""
    hi = dk[..., 1:, 1:, :] - dk[..., 1:, :-1, :]
    lo = dk[..., :-1, :-1, :] - dk[..., :-1, 1:, :]
    gh = (S[..., None] * hi).sum(dim=-2)          # (...,M-1,D) -> point i+1
    gl = (S[..., None] * lo).sum(dim=-2)          # (...,M-1,D) -> point i
    z = torch.zeros_like(gh[..., :1, :])
    return torch.cat([z, gh], dim=-2) + torch.cat([gl, z], dim=-2)


def gram_grad_points_analytic(X, Y, static_kernel, dyadic_order, sym=False, naive=False, backend="c"):
    A, B, M, N, D = X.shape[0], Y.shape[0], X.shape[1], Y.shape[1], X.shape[2]
    U, Ks = gram_grid(X, Y, static_kernel, dyadic_order, sym, naive, backend)
    inc = increments(Ks, dyadic_order)
    U_rev = torch.from_numpy(solve_gram(torch.flip(inc, dims=[2, 3]).numpy(), sym, naive, backend))
    U_rev = torch.flip(U_rev, dims=[2, 3])
    S = coarse_sensitivity(U, U_rev, dyadic_order)
    kind = "rbf" if isinstance(static_kernel, RBFKernel) else "linear_gram"
    dk = _d1k(static_kernel, kind, X[:, None, :, None, :], Y[None, :, None, :, :], Ks)
    return U[:, :, -1, -1], grad_points_from_S(S, dk), S


def batch_grad_points_analytic(X, Y, static_kernel, dyadic_order, naive=False, backend="c"):
    U, Ks = batch_grid(X, Y, static_kernel, dyadic_order, naive, backend)
    inc = increments(Ks, dyadic_order)
    U_rev = torch.from_numpy(solve_batch(torch.flip(inc, dims=[1, 2]).numpy(), naive, backend))
    U_rev = torch.flip(U_rev, dims=[1, 2])
    S = coarse_sensitivity(U, U_rev, dyadic_order)
    kind = "rbf" if isinstance(static_kernel, RBFKernel) else "linear_batch"
    dk = _d1k(static_kernel, kind, X[:, :, None, :], Y[:, None, :, :], Ks)
    return U[:, -1, -1], grad_points_from_S(S, dk), S


# ---------------------------------------------------------------------------------------
# kernel + first / second directional derivative along gamma
# (sigkernel.py:504-593 k_kgrad; solver = cuda_backend.py:165-223, see solver.c)
# ---------------------------------------------------------------------------------------
def derivative_increments(X, Y, gamma, static_kernel, dyadic_order, eps=1e-4):
    """The three refined increment tensors of k_kgrad (sigkernel.py:524-544), operation by operation."""
    G = static_kernel.Gram_matrix(X, Y)
    inc = second_difference(G)
    d1 = -(1. / eps) * G
    d2 = (1. / eps) * static_kernel.Gram_matrix(X + eps * gamma, Y)
    inc_d = second_difference(d1) + second_difference(d2)
    dd1 = -(1. / eps) * d1
    dd2 = -(2. / eps) * d2
    dd3 = (1. / eps ** 2) * static_kernel.Gram_matrix(X + 2. * eps * gamma, Y)
    inc_dd = second_difference(dd1) + second_difference(dd2) + second_difference(dd3)
    return refine(inc, dyadic_order), refine(inc_d, dyadic_order), refine(inc_dd, dyadic_order)


def solve_derivatives(inc, inc_d, inc_dd):
    """(A,B,MM,NN) x 3 -> K, K_diff, K_diffdiff, each (A,B)."""
    A, B, MM, NN = inc.shape
    arrs = [np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float64) for t in (inc, inc_d, inc_dd)]
    out = np.empty((A * B, 3), dtype=np.float64)
    work = np.empty(3 * (MM + 1) * (NN + 1), dtype=np.float64)
    _c().skb_oracle_solve_derivatives(_dptr(arrs[0]), _dptr(arrs[1]), _dptr(arrs[2]), A * B, MM, NN,
                                      _dptr(out), _dptr(work))
    out = torch.from_numpy(out).reshape(A, B, 3)
    return out[..., 0].clone(), out[..., 1].clone(), out[..., 2].clone()


def compute_kernel_and_derivatives_Gram(X, Y, gamma, static_kernel, dyadic_order, eps=1e-4):
    """SigKernel.compute_kernel_and_derivatives_Gram (sigkernel.py:43-89) without the max_batch splitting."""
    return solve_derivatives(*derivative_increments(X, Y, gamma, static_kernel, dyadic_order, eps))
