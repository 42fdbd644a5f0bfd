"""User API: a drop-in for the reference's `sigkernel.SigKernel` (sigkernel/sigkernel.py:15-197)
and its two autograd operators `_SigKernel` / `_SigKernelGram` (sigkernel.py:201-416), with the same
names, positional order, defaults, output shapes and assertion behaviour -- but every solve goes to
the hand-written sm_100a kernels behind include/sigkernel_b200.h.  CUDA tensors only.

Differences that are deliberate:
  * `max_batch` is accepted and ignored: the fused kernels never materialise a per-pair grid, so the
    reference's recursive halving (sigkernel.py:31-39, 102-127) has nothing to bound.
  * there is no limit on max(MM, NN) (reference asserts < 1024, sigkernel.py:222, 368).
  * gradients use the analytic derivative of the built-in static kernels instead of the reference's
    h = 1e-9 finite difference; plugin kernels keep the finite-difference route.
"""
import math

import torch

from . import ops

_H_FD = 1e-9   # finite-difference step the reference uses for d(static kernel)/dx (sigkernel.py:314, 473)


def _fused(static_kernel, gram):
    spec = getattr(static_kernel, "fused_spec", None)
    return spec(gram) if spec is not None else None


def _second_diff_x(Kh, Ks):
    """Finite-difference d inc_c / d x from the perturbed static matrix Kh (...,M,N,D) and Ks (...,M,N):
    returns (hi, lo) with hi[i,j] ~ h * d inc_c[i,j] / d x_{i+1}, lo[i,j] ~ h * d inc_c[i,j] / d x_i
    (sigkernel.py:483-487: Diff_1 = hi, Diff_2 = hi + lo)."""
    K = Ks[..., None]
    hi = Kh[..., 1:, 1:, :] - Kh[..., 1:, :-1, :] - K[..., 1:, 1:, :] + K[..., 1:, :-1, :]
    lo = -Kh[..., :-1, 1:, :] + Kh[..., :-1, :-1, :] + K[..., :-1, 1:, :] - K[..., :-1, :-1, :]
    return hi, lo


def _grad_points_from_sensitivity(S, hi, lo):
    """S (...,M-1,N-1), hi/lo (...,M-1,N-1,D) -> per-point gradient (...,M,D): coarse cell (i,j) feeds
    point i+1 through hi and point i through lo (sigkernel.py:489-500 collapsed to coarse cells)."""
    gh = (S[..., None] * hi).sum(dim=-2)
    gl = (S[..., None] * lo).sum(dim=-2)
    z = torch.zeros_like(gh[..., :1, :])
    return torch.cat([z, gh], dim=-2) + torch.cat([gl, z], dim=-2)


def _perturbed(X):
    """(A,M,D) -> (A, M*D, D): row (p,c) is x_p + h e_c  (sigkernel.py:316-318, 475-477)."""
    A, M, D = X.shape
    eye = torch.eye(D, dtype=X.dtype, device=X.device)
    return (X[:, :, None, :] + _H_FD * eye[None, None]).reshape(A, M * D, D)


class _SigKernel(torch.autograd.Function):
    """k(X^a, Y^a), a = 1..batch.  Same `apply` signature as the reference (sigkernel.py:204)."""

    @staticmethod
    def forward(ctx, X, Y, static_kernel, dyadic_order, _naive_solver=False):
        spec = _fused(static_kernel, gram=False)
        need_grad = X.requires_grad
        if spec is not None:
            kind, param, _ = spec
            if need_grad:
                K, gp = ops.sigkernel_forward_backward(X, Y, kind, param, dyadic_order, "batch", _naive_solver)
            else:
                K = ops.sigkernel_forward(X, Y, kind, param, dyadic_order, "batch", _naive_solver)
        else:
            Ks = static_kernel.batch_kernel(X, Y)
            if need_grad:
                K, S = ops.sensitivity_from_static(Ks, dyadic_order, "batch", _naive_solver)
                A, M, D = X.shape
                Kh = static_kernel.batch_kernel(_perturbed(X), Y).reshape(A, M, D, -1).permute(0, 1, 3, 2)
                hi, lo = _second_diff_x(Kh.to(torch.float64), Ks.to(torch.float64))
                gp = _grad_points_from_sensitivity(S, hi, lo) / _H_FD
            else:
                K = ops.sigkernel_forward_from_static(Ks, dyadic_order, "batch", _naive_solver)
        if need_grad:
            ctx.save_for_backward(gp)
        ctx.in_dtype = X.dtype
        return K.to(X.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        (gp,) = ctx.saved_tensors
        grad = grad_output.to(gp.dtype)[:, None, None] * gp
        return grad.to(ctx.in_dtype), None, None, None, None


class _SigKernelGram(torch.autograd.Function):
    """k(X^a, Y^b) for all a, b.  Same `apply` signature as the reference (sigkernel.py:349)."""

    @staticmethod
    def forward(ctx, X, Y, static_kernel, dyadic_order, sym=False, _naive_solver=False):
        spec = _fused(static_kernel, gram=True)
        need_grad = X.requires_grad
        pairs = "sym" if sym else "gram"
        if spec is not None:
            kind, param, _ = spec
            if need_grad:
                # like the reference, the whole backward is computed eagerly (sigkernel.py:397-399)
                # (a symmetric Gram still needs d k(X_a, X_b) / d X_a for every ordered pair)
                G, gp = ops.sigkernel_forward_backward(X, Y, kind, param, dyadic_order, "gram", _naive_solver)
            else:
                G = ops.sigkernel_forward(X, Y, kind, param, dyadic_order, pairs, _naive_solver)
        else:
            Ks = static_kernel.Gram_matrix(X, Y)
            if need_grad:
                G, S = ops.sensitivity_from_static(Ks, dyadic_order, "gram", _naive_solver)
                A, M, D = X.shape
                B = Y.shape[0]
                Kh = static_kernel.Gram_matrix(_perturbed(X), Y).reshape(A, B, M, D, -1).permute(0, 1, 2, 4, 3)
                hi, lo = _second_diff_x(Kh.to(torch.float64), Ks.to(torch.float64))
                gp = _grad_points_from_sensitivity(S, hi, lo) / _H_FD
            else:
                G = ops.sigkernel_forward_from_static(Ks, dyadic_order, pairs, _naive_solver)
        if need_grad:
            ctx.save_for_backward(gp)
            # the reference doubles the gradient when Y requires grad too, i.e. when Y is X and the
            # upstream gradient is symmetric (sigkernel.py:410-412); it never returns a gradient for Y
            ctx.double = bool(Y.requires_grad)
        ctx.in_dtype = X.dtype
        return G.to(X.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        (gp,) = ctx.saved_tensors
        grad = torch.einsum('ab,abmd->amd', grad_output.to(gp.dtype), gp)
        if ctx.double:
            grad = 2 * grad
        return grad.to(ctx.in_dtype), None, None, None, None, None


class _NoGradCtx:
    """Stand-in for the autograd context when nothing requires grad: the operators' forward is called directly
    (the result is the same tensor `apply` would return; skipping the autograd.Function machinery saves ~10 us per
    call, which matters for small batches)."""

    def save_for_backward(self, *tensors):
        pass


def _prepare(static_kernel, X, Y, gram):
    """Apply the path transform of a function-space kernel (differentiable torch ops) so that the
    autograd operators always see (batch, length, dim) paths."""
    spec = _fused(static_kernel, gram)
    if spec is not None and spec[2] is not None:
        return spec[2](X), spec[2](Y)
    return X, Y


def _offdiag_mean(K):
    n = K.shape[0]
    return (torch.sum(K) - torch.sum(torch.diag(K))) / (n * (n - 1.))


class SigKernel:
    """Signature kernel k_sig(x,y) = <S(f(x)), S(f(y))> for a static kernel k(x,y) = <f(x), f(y)>.
    Constructor and methods as in the reference (sigkernel.py:15-197)."""

    def __init__(self, static_kernel, dyadic_order, _naive_solver=False):
        self.static_kernel = static_kernel
        self.dyadic_order = dyadic_order
        self._naive_solver = _naive_solver

    def compute_kernel(self, X, Y, max_batch=100):
        """X (batch, len_x, dim), Y (batch, len_y, dim) -> (batch,)."""
        X, Y = _prepare(self.static_kernel, X, Y, gram=False)
        if not (X.requires_grad or Y.requires_grad):
            return _SigKernel.forward(_NoGradCtx(), X, Y, self.static_kernel, self.dyadic_order, self._naive_solver)
        return _SigKernel.apply(X, Y, self.static_kernel, self.dyadic_order, self._naive_solver)

    def compute_Gram(self, X, Y, sym=False, max_batch=100):
        """X (batch_x, len_x, dim), Y (batch_y, len_y, dim) -> (batch_x, batch_y)."""
        X, Y = _prepare(self.static_kernel, X, Y, gram=True)
        if not (X.requires_grad or Y.requires_grad):
            return _SigKernelGram.forward(_NoGradCtx(), X, Y, self.static_kernel, self.dyadic_order, sym, self._naive_solver)
        return _SigKernelGram.apply(X, Y, self.static_kernel, self.dyadic_order, sym, self._naive_solver)

    def compute_kernel_and_derivatives_Gram(self, X, Y, gamma, max_batch=100):
        """X (batch_x, len_x, dim), Y (batch_y, len_y, dim), gamma (batch_x, len_x, dim) ->
        k(X^i, Y^j), its directional derivative along gamma^i and the second one, each (batch_x, batch_y)
        (reference sigkernel.py:43-89 / k_kgrad :504-593).  No gradients flow through this call, as in the
        reference (every tensor is detached there before the solve)."""
        return k_kgrad(X, Y, gamma, self.dyadic_order, self.static_kernel)

    def compute_distance(self, X, Y, max_batch=100):
        """mean_a ||S(X^a) - S(Y^a)||^2."""
        assert not Y.requires_grad, "the second input should not require grad"
        K_XX = self.compute_kernel(X, X, max_batch)
        K_YY = self.compute_kernel(Y, Y, max_batch)
        K_XY = self.compute_kernel(X, Y, max_batch)
        return torch.mean(K_XX) + torch.mean(K_YY) - 2. * torch.mean(K_XY)

    def compute_scoring_rule(self, X, y, max_batch=100):
        """S(X, y) = E[k(X,X')] - 2 E[k(X,y)], y of shape (1, len_y, dim)."""
        assert not y.requires_grad, "the second input should not require grad"
        K_XX = self.compute_Gram(X, X, sym=True, max_batch=max_batch)
        K_Xy = self.compute_Gram(X, y, sym=False, max_batch=max_batch)
        return _offdiag_mean(K_XX) - 2. * torch.mean(K_Xy)

    def compute_expected_scoring_rule(self, X, Y, max_batch=100):
        """E_y[S(X, y)] over the sample Y."""
        assert not Y.requires_grad, "the second input should not require grad"
        K_XX = self.compute_Gram(X, X, sym=True, max_batch=max_batch)
        K_XY = self.compute_Gram(X, Y, sym=False, max_batch=max_batch)
        return _offdiag_mean(K_XX) - 2. * torch.mean(K_XY)

    def compute_mmd(self, X, Y, max_batch=100):
        """Unbiased MMD^2 between the samples X and Y."""
        assert not Y.requires_grad, "the second input should not require grad"
        K_XX = self.compute_Gram(X, X, sym=True, max_batch=max_batch)
        K_YY = self.compute_Gram(Y, Y, sym=True, max_batch=max_batch)
        K_XY = self.compute_Gram(X, Y, sym=False, max_batch=max_batch)
        return _offdiag_mean(K_XX) + _offdiag_mean(K_YY) - 2. * torch.mean(K_XY)


def k_kgrad(X, Y, gamma, dyadic_order, static_kernel, eps=1e-4):
    """Signature kernel and its first / second directional derivatives along gamma (reference
    sigkernel.py:504-593).  The static kernel is evaluated three times through the plugin interface, as the
    reference does (:524-539); the finite differences in eps, the second differences, the dyadic refinement
    and the three coupled PDE stencils (cuda_backend.py:165-223) run in one CUDA kernel pair."""
    with torch.no_grad():
        K0 = static_kernel.Gram_matrix(X, Y)
        K1 = static_kernel.Gram_matrix(X + eps * gamma, Y)
        K2 = static_kernel.Gram_matrix(X + 2. * eps * gamma, Y)
        K, Kd, Kdd = ops.kernel_and_derivatives_from_static(K0, K1, K2, dyadic_order, eps)
    return K.to(X.dtype), Kd.to(X.dtype), Kdd.to(X.dtype)


def c_alpha(m, alpha):
    return 4. * math.sqrt(-math.log(alpha) / m)


def hypothesis_test(y_pred, y_test, static_kernel, confidence_level=0.99, dyadic_order=0):
    """Two-sample MMD test (reference sigkernel.py:624-641); prints the verdict and also returns it."""
    TU = SigKernel(static_kernel, dyadic_order).compute_mmd(y_pred, y_test)
    rejected = bool(TU > c_alpha(max(y_pred.shape[0], y_test.shape[0]), confidence_level))
    verdict = "rejected: distribution are not equal" if rejected else "accepted: distribution are equal"
    print(f'Hypothesis {verdict} with {confidence_level*100}% confidence')
    return rejected


def SigCHSIC(X, Y, Z, static_kernel, dyadic_order=1, eps=0.1):
    """Signature conditional HSIC of (X, Y) given Z from three symmetric Gram matrices
    (reference sigkernel.py:644-691)."""
    m = X.shape[0]
    sk = SigKernel(static_kernel, dyadic_order)
    H = torch.eye(m, dtype=X.dtype, device=X.device) - 1. / m
    KX, KY, KZ = (H @ sk.compute_Gram(T, T, sym=True) @ H for T in (X, Y, Z))
    KZe_inv = torch.cholesky_inverse(KZ + m * eps * torch.eye(m, dtype=X.dtype, device=X.device))
    Amat = KZ @ (KZe_inv @ KZe_inv) @ KZ
    Bmat = KX @ Amat @ KY
    return (torch.trace(KX @ KY) - 2. * torch.trace(Bmat) + torch.trace(Bmat @ Amat)) / m ** 2
