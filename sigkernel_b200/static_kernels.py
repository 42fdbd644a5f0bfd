"""Static-kernel plugin surface (mirror of the reference's sigkernel/static_kernels.py).

A static kernel is any object with
    batch_kernel(X, Y) -> (A, M, N)        k(X^a_s, Y^a_t)
    Gram_matrix(X, Y)  -> (A, B, M, N)     k(X^a_s, Y^b_t)
(reference static_kernels.py:17-33).  Those two methods stay callable from Python, on any device,
exactly as in the reference -- they ARE the plugin interface.  What is new: a kernel may also
implement
    fused_spec(gram: bool) -> (kind, param, transform)
telling SigKernel that the CUDA solver can evaluate it on the fly ("linear": param*<x,y>,
"rbf": exp(-|x-y|^2/param)) after applying `transform` to each path tensor.  Only the EXACT built-in
types advertise it (a subclass overriding batch_kernel must not silently get the fused arithmetic);
everything else goes through Gram_matrix/batch_kernel and the solver's from-static entry point.
"""
import math

import torch


def _flat(X):
    """(batch, len_t, len_x, dim) function-valued paths -> (batch, len_t, len_x*dim)."""
    return X.reshape(X.shape[0], X.shape[1], -1)


class LinearKernel:
    """k(x,y) = <x,y>.  Reference static_kernels.py:11-33, including its inconsistency: the batch
    form scales both arguments by `scale` (:24), the Gram form ignores `scale` (:33)."""

    def __init__(self, scale=1.0):
        self.scale = scale

    def batch_kernel(self, X, Y):
        return torch.bmm(self.scale * X, (self.scale * Y).transpose(1, 2))

    def Gram_matrix(self, X, Y):
        return torch.einsum('ipk,jqk->ijpq', X, Y)

    def fused_spec(self, gram):
        if type(self) is not LinearKernel:
            return None
        return "linear", (1.0 if gram else float(self.scale) ** 2), None


class RBFKernel:
    """k(x,y) = exp(-|x-y|^2 / sigma)  (sigma, not 2 sigma^2: reference static_kernels.py:36-73)."""

    def __init__(self, sigma):
        self.sigma = sigma

    @staticmethod
    def _sqdist(xy, xs, ys):
        return (-2. * xy) + (xs + ys)

    def batch_kernel(self, X, Y):
        xs = (X ** 2).sum(dim=2)[:, :, None]
        ys = (Y ** 2).sum(dim=2)[:, None, :]
        return torch.exp(-self._sqdist(torch.bmm(X, Y.transpose(1, 2)), xs, ys) / self.sigma)

    def Gram_matrix(self, X, Y):
        xs = (X ** 2).sum(dim=2)[:, None, :, None]
        ys = (Y ** 2).sum(dim=2)[None, :, None, :]
        return torch.exp(-self._sqdist(torch.einsum('ipk,jqk->ijpq', X, Y), xs, ys) / self.sigma)

    def fused_spec(self, gram):
        if type(self) is not RBFKernel:
            return None
        return "rbf", float(self.sigma), None


# ---- function-space kernels: paths are (batch, len_t, len_x, dim) ------------------------------
def cos_exp_kernel(x_y, n_freqs=5, sigma=1):
    """sum_{n<n_freqs} cos(2 pi n (x-y)) * exp(-(x-y)^2 / sigma)   (reference static_kernels.py:233-250)."""
    n = torch.arange(n_freqs, device=x_y.device)
    return torch.cos(2 * math.pi * x_y[..., None] * n).sum(dim=-1) * torch.exp(-x_y ** 2 / sigma)


def CEXP(X, n_freqs=20, sigma=math.sqrt(10)):
    """Integral operator of the cos-exp kernel applied along the len_x axis of X
    (batch, len_t, len_x, dim), functions sampled on a uniform grid of [0,1]
    (reference static_kernels.py:208-231)."""
    L = X.shape[2]
    grid = torch.linspace(0, 1, L, dtype=torch.float64, device=X.device)
    T = cos_exp_kernel(grid[:, None] - grid[None, :], n_freqs=n_freqs, sigma=sigma)
    return torch.einsum('btxd,xy->btyd', X, T.to(X.dtype)) / L


class Linear_ID_Kernel(LinearKernel):
    """Linear kernel on flattened function values (reference static_kernels.py:146-175)."""

    def __init__(self):
        super().__init__()

    def batch_kernel(self, X, Y):
        return super().batch_kernel(_flat(X), _flat(Y))

    def Gram_matrix(self, X, Y):
        return super().Gram_matrix(_flat(X), _flat(Y))

    def fused_spec(self, gram):
        if type(self) is not Linear_ID_Kernel:
            return None
        return "linear", (1.0 if gram else float(self.scale) ** 2), _flat


class RBF_ID_Kernel(RBFKernel):
    """RBF kernel on flattened function values (reference static_kernels.py:178-206)."""

    def __init__(self, sigma):
        super().__init__(sigma)

    def batch_kernel(self, X, Y):
        return super().batch_kernel(_flat(X), _flat(Y))

    def Gram_matrix(self, X, Y):
        return super().Gram_matrix(_flat(X), _flat(Y))

    def fused_spec(self, gram):
        if type(self) is not RBF_ID_Kernel:
            return None
        return "rbf", float(self.sigma), _flat


class RBF_CEXP_Kernel(RBFKernel):
    """RBF kernel after the cos-exp integral transform (reference static_kernels.py:75-115)."""

    def __init__(self, sigma1, sigma2, n_freqs):
        self.sigma1 = sigma1
        super().__init__(sigma2)
        self.n_freqs = n_freqs

    def _transform(self, X):
        return _flat(CEXP(X, self.n_freqs, self.sigma1))

    def batch_kernel(self, X, Y):
        return super().batch_kernel(self._transform(X), self._transform(Y))

    def Gram_matrix(self, X, Y):
        return super().Gram_matrix(self._transform(X), self._transform(Y))

    def fused_spec(self, gram):
        if type(self) is not RBF_CEXP_Kernel:
            return None
        return "rbf", float(self.sigma), self._transform


class RBF_SQR_Kernel:
    """Product of an RBF kernel on the values and one on their squares.  (The reference constructor,
    static_kernels.py:117-122, raises NameError -- `sigma_1` undefined; this one works.)  Not fusable:
    served through the from-static plugin path."""

    def __init__(self, sigma1, sigma2):
        self.rbf1 = RBFKernel(sigma1)
        self.rbf2 = RBFKernel(sigma2)

    def batch_kernel(self, X, Y):
        X, Y = _flat(X), _flat(Y)
        return self.rbf1.batch_kernel(X, Y) * self.rbf2.batch_kernel(X ** 2, Y ** 2)

    def Gram_matrix(self, X, Y):
        X, Y = _flat(X), _flat(Y)
        return self.rbf1.Gram_matrix(X, Y) * self.rbf2.Gram_matrix(X ** 2, Y ** 2)
