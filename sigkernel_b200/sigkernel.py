"""User API: a drop-in for the reference's `sigkernel.SigKernel` (sigkernel/sigkernel.py:15-197)
and its two autograd operators `_SigKernel` / `_SigKernelGram` (sigkernel.py:201-416), with the same
names, positional order, defaults, output shapes and assertion behaviour -- but every solve goes to
the hand-written sm_100a kernels behind include/sigkernel_b200.h.  CUDA tensors only.

Differences that are deliberate:
  * `max_batch` only matters for plugin static kernels (their (A,B,M,N) matrices are bounded by solving blocks of at
    most max_batch x max_batch paths); the fused kernels never materialise anything per pair, so the reference's
    recursive halving (sigkernel.py:31-39, 102-127) has nothing to bound there.
  * the backward of the built-in static kernels is lazy (it runs when autograd asks for it, contracted on the fly with
    d loss / d K) where the reference computes grad_points eagerly inside forward (sigkernel.py:397-399); under
    torch.no_grad() nothing of the backward is computed.
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


def _same_paths(X, Y):
    """True if Y is (a view of) the very tensor X: the precondition of the reference's sym=True shortcut
    (cython_backend.pyx:76-97 solves a <= b and mirrors, which is only right for X == Y)."""
    return X is Y or (X.shape == Y.shape and X.dtype == Y.dtype and X.device == Y.device and
                      X.data_ptr() == Y.data_ptr() and X.stride() == Y.stride())


def _chunks(n, max_batch):
    """Row blocks of at most max_batch (the reference bounds its per-pair tensors by recursive halving,
    sigkernel.py:31-39, 102-127; equal-size blocks do the same job)."""
    max_batch = max(1, int(max_batch))
    k = -(-n // max_batch)
    size = -(-n // k)
    return [(lo, min(n, lo + size)) for lo in range(0, n, size)]


class _SigKernel(torch.autograd.Function):
    """k(X^a, Y^a), a = 1..batch.  Same `apply` signature as the reference (sigkernel.py:204).  Forward and backward of
    the built-in static kernels: skb_sigkernel_fwd_ctx / skb_sigkernel_bwd_vjp (the backward runs when autograd asks for
    it, fused with the contraction by grad_output); shapes outside those kernels and plugin kernels compute the
    reference's grad_points eagerly, as the reference does (sigkernel.py:256-343)."""

    @staticmethod
    def forward(ctx, X, Y, static_kernel, dyadic_order, _naive_solver=False):
        spec = _fused(static_kernel, gram=False)
        need_grad = X.requires_grad
        ctx.mode = None
        if spec is not None:
            kind, param, _ = spec
            if need_grad:
                res = ops.sigkernel_forward_ctx(X, Y, kind, param, dyadic_order, "batch", _naive_solver)
                if res is not None:
                    K, bctx = res
                    ctx.mode = "lazy"
                    ctx.meta = (kind, param, dyadic_order, _naive_solver)
                    ctx.save_for_backward(X, Y, bctx)
                else:
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
        if need_grad and ctx.mode is None:
            ctx.mode = "eager"
            ctx.save_for_backward(gp)
        ctx.in_dtype = X.dtype
        return K.to(X.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.mode is None:
            # only Y required grad: the reference never returns a gradient for Y (sigkernel.py:343)
            return None, None, None, None, None
        if ctx.mode == "lazy":
            X, Y, bctx = ctx.saved_tensors
            kind, param, d, naive = ctx.meta
            grad = ops.sigkernel_backward_vjp(X, Y, kind, param, d, "batch", bctx, "batch", grad_out=grad_output, naive=naive)
        else:
            (gp,) = ctx.saved_tensors
            grad = grad_output.to(gp.dtype)[:, None, None] * gp
        return grad.to(ctx.in_dtype), None, None, None, None


class _SigKernelGram(torch.autograd.Function):
    """k(X^a, Y^b) for all a, b.  Same `apply` signature as the reference (sigkernel.py:349)."""

    @staticmethod
    def forward(ctx, X, Y, static_kernel, dyadic_order, sym=False, _naive_solver=False):
        spec = _fused(static_kernel, gram=True)
        need_grad = X.requires_grad
        sym = bool(sym) and _same_paths(X, Y)     # the triangular shortcut is only right for X == Y
        pairs = "sym" if sym else "gram"
        ctx.mode = None
        if spec is not None:
            kind, param, _ = spec
            if need_grad:
                # with gradients the full square is solved even when symmetric: the reversed sweep runs over every
                # ordered pair and rebuilds each grid from ITS OWN last row / column (the transposed boundaries of (b, a)
                # differ from those of (a, b) in the last bits, which the backward recurrence amplifies past the check)
                # (unless the unordered-pair sweep covers the shape: it runs the pair (a, b), a <= b, once, from its own
                #  boundaries, and gets the term of (b, a) out of the same sensitivities)
                gpairs = "sym" if sym and ops.adjoint_sym_supported(X.shape[1], X.shape[2], dyadic_order, kind, _naive_solver) else "gram"
                res = ops.sigkernel_forward_ctx(X, Y, kind, param, dyadic_order, gpairs, _naive_solver)
                if res is not None:
                    G, bctx = res
                    ctx.mode = "lazy"
                    ctx.meta = (kind, param, dyadic_order, _naive_solver, gpairs)
                    ctx.save_for_backward(X, Y, bctx)
                else:
                    # like the reference, the whole backward is computed eagerly (sigkernel.py:397-399)
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
        if need_grad and ctx.mode is None:
            ctx.mode = "eager"
            ctx.save_for_backward(gp)
        # the reference doubles the gradient when Y requires grad too, i.e. when Y is X and the
        # upstream gradient is symmetric (sigkernel.py:410-412); it never returns a gradient for Y
        ctx.double = bool(Y.requires_grad)
        ctx.in_dtype = X.dtype
        return G.to(X.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.mode is None:
            return None, None, None, None, None, None
        scale = 2.0 if ctx.double else 1.0
        if ctx.mode == "lazy":
            X, Y, bctx = ctx.saved_tensors
            kind, param, d, naive, pairs = ctx.meta
            grad = ops.sigkernel_backward_vjp(X, Y, kind, param, d, pairs, bctx, pairs, grad_out=grad_output,
                                              out_scale=scale, naive=naive)
        else:
            (gp,) = ctx.saved_tensors
            grad = torch.einsum('ab,abmd->amd', grad_output.to(gp.dtype), gp)
            if ctx.double:
                grad = 2 * grad
        return grad.to(ctx.in_dtype), None, None, None, None, None


class _SigLoss(torch.autograd.Function):
    """Loss heads of the reference built from weighted sums of Gram entries (sigkernel.py:130-197):
        mmd       mean_offdiag k(X,X) + mean_offdiag k(Y,Y) - 2 mean k(X,Y)
        score     mean_offdiag k(X,X) - 2 mean k(X,Y)                 (compute_scoring_rule / expected variant)
        distance  mean k(X^a,X^a) + mean k(Y^a,Y^a) - 2 mean k(X^a,Y^a)
    Forward: one solver launch per Gram plus one reduction launch each; backward: one reversed sweep per Gram that involves
    X, contracted on the fly with the closed-form d loss / d K (a constant off the diagonal, another on it) -- neither the
    (A,B) upstream-gradient matrices nor the (A,B,M,D) tensor of sigkernel.py:405-416 exist, and no torch kernel runs.
    Only for the built-in static kernels on shapes the reconstruction adjoint covers; anything else composes
    compute_Gram / compute_kernel as the reference does."""

    @staticmethod
    def terms(which, n, m):
        # (first, second, pairs, w_diag, w_off, gradient scale or None)
        if which == "mmd":
            return [("X", "X", "sym", 0.0, 1.0 / (n * (n - 1.0)), 2.0), ("Y", "Y", "sym", 0.0, 1.0 / (m * (m - 1.0)), None),
                    ("X", "Y", "gram", -2.0 / (n * m), -2.0 / (n * m), 1.0)]
        if which == "score":
            return [("X", "X", "sym", 0.0, 1.0 / (n * (n - 1.0)), 2.0), ("X", "Y", "gram", -2.0 / (n * m), -2.0 / (n * m), 1.0)]
        if which == "distance":
            return [("X", "X", "batch", 1.0 / n, 0.0, 1.0), ("Y", "Y", "batch", 1.0 / n, 0.0, None),
                    ("X", "Y", "batch", -2.0 / n, 0.0, 1.0)]
        raise ValueError(which)

    @staticmethod
    def supported(X, Y, static_kernel, dyadic_order, naive, which):
        spec = _fused(static_kernel, gram=which != "distance")
        if spec is None or not (X.is_cuda and Y.is_cuda) or X.dim() != 3 or Y.dim() != 3:
            return False
        if which != "distance" and (X.shape[0] < 2 or (which == "mmd" and Y.shape[0] < 2)):
            return False
        kind = spec[0]
        D = X.shape[2]
        return all(ops.adjoint_plan(P.shape[1], Q.shape[1], D, dyadic_order, kind, naive) == 6
                   for P, Q in ((X, X), (X, Y)))

    @staticmethod
    def forward(ctx, X, Y, static_kernel, dyadic_order, _naive_solver, which):
        kind, param, _ = _fused(static_kernel, gram=which != "distance")
        T = {"X": X, "Y": Y}
        need_grad = X.requires_grad
        acc = None
        saved, plan = [], []
        for first, second, pairs, w_diag, w_off, gscale in _SigLoss.terms(which, X.shape[0], Y.shape[0]):
            P, Q = T[first], T[second]
            if need_grad and gscale is not None:
                if pairs == "sym":
                    if not ops.adjoint_sym_supported(P.shape[1], P.shape[2], dyadic_order, kind, _naive_solver):
                        pairs = "gram"    # no unordered-pair sweep for this shape: the full square is solved
                G, bctx = ops.sigkernel_forward_ctx(P, Q, kind, param, dyadic_order, pairs, _naive_solver)
                plan.append((second, pairs, w_diag, w_off, gscale, len(saved)))
                saved.append(bctx)
            else:
                G = ops.sigkernel_forward(P, Q, kind, param, dyadic_order, pairs, _naive_solver)
            acc = ops.gram_weighted_sum(G, pairs, w_diag, w_off, acc)
        ctx.plan = plan
        ctx.meta = (kind, param, dyadic_order, _naive_solver)
        ctx.in_dtype = X.dtype
        if need_grad:
            ctx.save_for_backward(X, Y, *saved)
        out = acc.view(())
        return out if X.dtype == torch.float64 else out.to(X.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        if not ctx.plan:
            return None, None, None, None, None, None
        X, Y, *saved = ctx.saved_tensors
        kind, param, d, naive = ctx.meta
        grad = None
        for second, pairs, w_diag, w_off, gscale, k in ctx.plan:
            Q = X if second == "X" else Y
            grad = ops.sigkernel_backward_vjp(X, Q, kind, param, d, pairs, saved[k], pairs,
                                              w_diag=w_diag, w_off=w_off, out_scale=gscale, out_scale_dev=grad_output,
                                              into=grad, naive=naive)
        return (grad if ctx.in_dtype == torch.float64 else grad.to(ctx.in_dtype)), None, None, None, None, None


class _NoGradCtx:
    """Stand-in for the autograd context when nothing requires grad: the operators' forward is called directly
    (the result is the same tensor `apply` would return; skipping the autograd.Function machinery saves ~10 us per
    call, which matters for small batches)."""
    mode = None

    def save_for_backward(self, *tensors):
        pass


def _prepare(static_kernel, X, Y, gram):
    """Apply the path transform of a function-space kernel (differentiable torch ops) so that the
    autograd operators always see (batch, length, dim) paths."""
    spec = _fused(static_kernel, gram)
    if spec is not None and spec[2] is not None:
        return spec[2](X), (spec[2](Y) if Y is not X else None)
    return X, Y


def _offdiag_mean(K):
    n = K.shape[0]
    return (torch.sum(K) - torch.sum(torch.diag(K))) / (n * (n - 1.))


def _wants_grad(*tensors):
    return torch.is_grad_enabled() and any(t.requires_grad for t in tensors)


class SigKernel:
    """Signature kernel k_sig(x,y) = <S(f(x)), S(f(y))> for a static kernel k(x,y) = <f(x), f(y)>.
    Constructor and methods as in the reference (sigkernel.py:15-197).

    max_batch: the fused static kernels (LinearKernel, RBFKernel and the function-space kernels built on them) never
    materialise anything per pair, so max_batch is ignored for them; plugin kernels go through Gram_matrix /
    batch_kernel and their (A,B,M,N) static matrices are bounded by solving blocks of at most max_batch x max_batch
    paths, like the reference's recursive halving (sigkernel.py:31-39, 102-127)."""

    def __init__(self, static_kernel, dyadic_order, _naive_solver=False):
        self.static_kernel = static_kernel
        self.dyadic_order = dyadic_order
        self._naive_solver = _naive_solver

    def _gram(self, X, Y, sym):
        if not _wants_grad(X, Y):
            return _SigKernelGram.forward(_NoGradCtx(), X, Y, self.static_kernel, self.dyadic_order, sym, self._naive_solver)
        return _SigKernelGram.apply(X, Y, self.static_kernel, self.dyadic_order, sym, self._naive_solver)

    def _batch(self, X, Y):
        if not _wants_grad(X, Y):
            return _SigKernel.forward(_NoGradCtx(), X, Y, self.static_kernel, self.dyadic_order, self._naive_solver)
        return _SigKernel.apply(X, Y, self.static_kernel, self.dyadic_order, self._naive_solver)

    def compute_kernel(self, X, Y, max_batch=100):
        """X (batch, len_x, dim), Y (batch, len_y, dim) -> (batch,)."""
        same = Y is X
        X, Y = _prepare(self.static_kernel, X, Y, gram=False)
        Y = X if (same and Y is None) else Y
        if _fused(self.static_kernel, gram=False) is not None or X.shape[0] <= max_batch:
            return self._batch(X, Y)
        return torch.cat([self._batch(X[lo:hi], Y[lo:hi]) for lo, hi in _chunks(X.shape[0], max_batch)], dim=0)

    def compute_Gram(self, X, Y, sym=False, max_batch=100):
        """X (batch_x, len_x, dim), Y (batch_y, len_y, dim) -> (batch_x, batch_y)."""
        same = Y is X
        X, Y = _prepare(self.static_kernel, X, Y, gram=True)
        Y = X if (same and Y is None) else Y
        if _fused(self.static_kernel, gram=True) is not None or (X.shape[0] <= max_batch and Y.shape[0] <= max_batch):
            return self._gram(X, Y, sym)
        # plugin kernels: blocks of at most max_batch x max_batch paths (sub-blocks are never symmetric, sigkernel.py:107-124)
        rows = []
        for lo, hi in _chunks(X.shape[0], max_batch):
            rows.append(torch.cat([self._gram(X[lo:hi], Y[l2:h2], False) for l2, h2 in _chunks(Y.shape[0], max_batch)], dim=1))
        return torch.cat(rows, dim=0)

    def compute_kernel_and_derivatives_Gram(self, X, Y, gamma, max_batch=100):
        """X (batch_x, len_x, dim), Y (batch_y, len_y, dim), gamma (batch_x, len_x, dim) ->
        k(X^i, Y^j), its directional derivative along gamma^i and the second one, each (batch_x, batch_y)
        (reference sigkernel.py:43-89 / k_kgrad :504-593).  No gradients flow through this call, as in the
        reference (every tensor is detached there before the solve).  Blocks of at most max_batch x max_batch paths
        bound the three (A,B,M,N) static matrices, as the reference's recursive halving does."""
        A, B = X.shape[0], Y.shape[0]
        if A <= max_batch and B <= max_batch:
            return k_kgrad(X, Y, gamma, self.dyadic_order, self.static_kernel)
        rows = [[], [], []]
        for lo, hi in _chunks(A, max_batch):
            parts = [k_kgrad(X[lo:hi], Y[l2:h2], gamma[lo:hi], self.dyadic_order, self.static_kernel) for l2, h2 in _chunks(B, max_batch)]
            for i in range(3):
                rows[i].append(torch.cat([p[i] for p in parts], dim=1))
        return tuple(torch.cat(r, dim=0) for r in rows)

    def _loss(self, X, Y, which):
        """Fused loss head when the static kernel and the shapes allow it, else None."""
        if _fused(self.static_kernel, gram=which != "distance") is None:
            return None
        Xp, Yp = _prepare(self.static_kernel, X, Y, gram=which != "distance")
        if Yp is None:
            Yp = Xp
        if not _SigLoss.supported(Xp, Yp, self.static_kernel, self.dyadic_order, self._naive_solver, which):
            return None
        if not _wants_grad(Xp, Yp):
            class _Ctx(_NoGradCtx):
                plan = None
            with torch.no_grad():
                return _SigLoss.forward(_Ctx(), Xp.detach(), Yp.detach(), self.static_kernel, self.dyadic_order, self._naive_solver, which)
        return _SigLoss.apply(Xp, Yp, self.static_kernel, self.dyadic_order, self._naive_solver, which)

    def compute_distance(self, X, Y, max_batch=100):
        """mean_a ||S(X^a) - S(Y^a)||^2."""
        assert not Y.requires_grad, "the second input should not require grad"
        fused = self._loss(X, Y, "distance") if X.shape[0] == Y.shape[0] else None
        if fused is not None:
            return fused
        K_XX = self.compute_kernel(X, X, max_batch)
        K_YY = self.compute_kernel(Y, Y, max_batch)
        K_XY = self.compute_kernel(X, Y, max_batch)
        return torch.mean(K_XX) + torch.mean(K_YY) - 2. * torch.mean(K_XY)

    def compute_scoring_rule(self, X, y, max_batch=100):
        """S(X, y) = E[k(X,X')] - 2 E[k(X,y)], y of shape (1, len_y, dim)."""
        assert not y.requires_grad, "the second input should not require grad"
        fused = self._loss(X, y, "score")
        if fused is not None:
            return fused
        K_XX = self.compute_Gram(X, X, sym=True, max_batch=max_batch)
        K_Xy = self.compute_Gram(X, y, sym=False, max_batch=max_batch)
        return _offdiag_mean(K_XX) - 2. * torch.mean(K_Xy)

    def compute_expected_scoring_rule(self, X, Y, max_batch=100):
        """E_y[S(X, y)] over the sample Y."""
        assert not Y.requires_grad, "the second input should not require grad"
        fused = self._loss(X, Y, "score")
        if fused is not None:
            return fused
        K_XX = self.compute_Gram(X, X, sym=True, max_batch=max_batch)
        K_XY = self.compute_Gram(X, Y, sym=False, max_batch=max_batch)
        return _offdiag_mean(K_XX) - 2. * torch.mean(K_XY)

    def compute_mmd(self, X, Y, max_batch=100):
        """Unbiased MMD^2 between the samples X and Y."""
        assert not Y.requires_grad, "the second input should not require grad"
        fused = self._loss(X, Y, "mmd")
        if fused is not None:
            return fused
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
        spec = _fused(static_kernel, gram=True)
        if spec is not None and spec[2] is None and X.is_cuda and X.dim() == 3:
            # the two built-in kernels: one CUDA pass per static matrix instead of the plugin's chain of torch ops
            gram = lambda x: ops.static_gram(x, Y, spec[0], spec[1], "gram")
        else:
            gram = lambda x: static_kernel.Gram_matrix(x, Y)
        K0 = gram(X)
        K1 = gram(X + eps * gamma)
        K2 = gram(X + 2. * eps * gamma)
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
