"""Generate tests/golden/*.npz by running the UNMODIFIED reference package on CPU fp64.

Run in the build container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

The reference tree holds no golden vectors of its own, so parity is pinned
to outputs of the reference itself: `sigkernel/*.py` imported from /root/reference with its
Cython solver compiled from `sigkernel/cython_backend.pyx` into oracle/_ref (top-level module
name `cython_backend`, as reference setup.py:46 and sigkernel.py:6 require).

Cases = BASELINE.json configs (cfg3-cfg5 sub-sampled: the first rows of the same seeded
tensors, so the stored block equals the corresponding block of the full Gram) + the case list
of the reference's only test file (sigkernel/test_mps.py:14-214) at fp64 + edge cases.
Every file stores the inputs, the parameters and the reference outputs.
"""
import glob
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, "/root/reference")
assert glob.glob(os.path.join(ROOT, "oracle", "_ref", "cython_backend*.so")), "run `make -C oracle ref`"

import sigkernel as ref  # noqa: E402  (the unmodified reference)


def make_kernel(spec):
    if spec[0] == "rbf":
        return ref.RBFKernel(sigma=spec[1])
    if spec[0] == "rbf_id":          # function-space kernels (static_kernels.py:75-206): 4-D inputs
        return ref.RBF_ID_Kernel(spec[1])
    if spec[0] == "linear_id":
        return ref.Linear_ID_Kernel()
    if spec[0] == "rbf_cexp":
        return ref.RBF_CEXP_Kernel(spec[1], spec[2], spec[3])
    return ref.LinearKernel(scale=spec[1])


def gen(kind, seed, shape_x, shape_y):
    g = torch.Generator().manual_seed(seed)
    if kind == "rand":
        return (torch.rand(shape_x, dtype=torch.float64, generator=g),
                torch.rand(shape_y, dtype=torch.float64, generator=g))
    if kind == "randn":
        return (torch.randn(shape_x, dtype=torch.float64, generator=g),
                torch.randn(shape_y, dtype=torch.float64, generator=g))
    if kind == "bm":   # Brownian-like smooth paths, cf. transformers.py:192-195
        X = torch.randn(shape_x, dtype=torch.float64, generator=g) / np.sqrt(shape_x[1])
        Y = torch.randn(shape_y, dtype=torch.float64, generator=g) / np.sqrt(shape_y[1])
        return torch.cumsum(X, 1), torch.cumsum(Y, 1)
    raise ValueError(kind)


# name, op, static kernel, dyadic order, naive, data kind, seed, X shape, Y shape, extra
CASES = [
    # BASELINE.json configs
    ("cfg1_kernel_linear", "kernel", ("linear", 1.0), 0, False, "rand", 0, (4, 10, 2), (4, 10, 2), {}),
    ("cfg2_gram_rbf", "gram", ("rbf", 0.5), 1, False, "rand", 0, (64, 32, 3), (64, 32, 3), {"sub": 16}),
    ("cfg3_gram_rbf", "gram", ("rbf", 0.5), 2, False, "rand", 0, (128, 64, 5), (128, 64, 5), {"sub": 8}),
    ("cfg4_mmd_bwd_rbf", "mmd_bwd", ("rbf", 0.5), 1, False, "rand", 0, (128, 64, 3), (128, 64, 3), {"sub": 6}),
    ("cfg5_gram_rbf", "gram", ("rbf", 0.5), 2, False, "rand", 0, (512, 128, 8), (512, 128, 8), {"sub": 3}),
    # test_mps.py case list, at fp64
    ("mps_basic_kernel", "kernel", ("rbf", 1.0), 0, False, "randn", 1, (3, 8, 2), (3, 10, 2), {}),
    ("mps_gram", "gram", ("rbf", 0.5), 0, False, "randn", 2, (4, 6, 2), (5, 6, 2), {}),
    ("mps_sym_gram", "gram_sym", ("rbf", 1.0), 0, False, "randn", 3, (5, 8, 2), (5, 8, 2), {}),
    ("mps_gradients", "kernel_bwd", ("rbf", 0.5), 0, False, "randn", 4, (3, 6, 2), (3, 6, 2), {}),
    ("mps_dyadic1", "kernel", ("rbf", 1.0), 1, False, "randn", 5, (2, 6, 2), (2, 6, 2), {}),
    ("mps_mmd", "mmd", ("rbf", 1.0), 0, False, "randn", 6, (4, 6, 2), (5, 6, 2), {}),
    ("mps_linear", "kernel", ("linear", 1.0), 0, False, "randn", 7, (3, 6, 2), (3, 6, 2), {}),
    ("mps_unequal", "gram", ("rbf", 1.0), 0, False, "randn", 8, (3, 5, 2), (4, 8, 2), {}),
    # extra coverage: schemes, dyadic orders, lengths that are not multiples of anything
    ("gram_rbf_naive", "gram", ("rbf", 0.5), 1, True, "rand", 9, (3, 9, 3), (4, 7, 3), {}),
    ("gram_linear_d3", "gram", ("linear", 1.0), 3, False, "bm", 10, (3, 11, 4), (2, 13, 4), {}),
    ("kernel_linear_scale", "kernel", ("linear", 0.5), 2, False, "bm", 11, (5, 12, 3), (5, 9, 3), {}),
    ("gram_rbf_long", "gram", ("rbf", 2.0), 1, False, "bm", 12, (2, 150, 2), (2, 97, 2), {}),
    ("gram_rbf_len2", "gram", ("rbf", 1.0), 2, False, "randn", 13, (3, 2, 2), (2, 2, 2), {}),
    ("gram_bwd_rbf", "gram_bwd", ("rbf", 0.5), 1, False, "rand", 14, (4, 9, 3), (5, 7, 3), {}),
    ("gram_bwd_rbf_d2", "gram_bwd", ("rbf", 1.5), 2, False, "bm", 15, (3, 8, 2), (3, 8, 2), {}),
    ("gram_bwd_linear", "gram_bwd", ("linear", 1.0), 1, False, "bm", 16, (3, 7, 3), (4, 10, 3), {}),
    ("gram_bwd_rbf_naive", "gram_bwd", ("rbf", 0.5), 0, True, "rand", 17, (3, 6, 2), (2, 6, 2), {}),
    ("kernel_bwd_linear_scale", "kernel_bwd", ("linear", 0.5), 1, False, "bm", 18, (4, 8, 3), (4, 6, 3), {}),
    ("gram_sym_bwd_rbf", "gram_sym_bwd", ("rbf", 0.5), 1, False, "rand", 19, (5, 8, 3), (5, 8, 3), {}),
    ("mmd_bwd_small", "mmd_bwd", ("rbf", 0.5), 1, False, "rand", 20, (5, 8, 3), (4, 8, 3), {}),
    ("scoring_rule", "scoring", ("rbf", 1.0), 1, False, "randn", 21, (5, 7, 2), (1, 7, 2), {}),
    ("distance", "distance", ("rbf", 1.0), 0, False, "randn", 22, (4, 7, 2), (4, 7, 2), {}),
    # function-space static kernels (static_kernels.py:75-206): paths are (batch, len_t, len_x, dim)
    ("fs_gram_rbf_id", "gram", ("rbf_id", 2.0), 1, False, "rand", 40, (3, 9, 4, 2), (4, 7, 4, 2), {}),
    ("fs_kernel_rbf_id", "kernel", ("rbf_id", 1.5), 0, False, "rand", 41, (3, 8, 3, 1), (3, 6, 3, 1), {}),
    ("fs_gram_linear_id", "gram", ("linear_id",), 2, False, "bm", 42, (3, 6, 2, 2), (2, 9, 2, 2), {}),
    ("fs_gram_rbf_cexp", "gram", ("rbf_cexp", 1.0, 2.0, 4), 1, False, "rand", 43, (3, 7, 4, 2), (3, 7, 4, 2), {}),
    ("fs_gram_rbf_id_wide", "gram", ("rbf_id", 8.0), 1, False, "rand", 44, (2, 6, 8, 2), (3, 6, 8, 2), {}),
    # (the reference's backward does not support 4-D inputs: prep_backward permutes a 5-D tensor with 4 indices,
    #  sigkernel.py:476 -> RuntimeError; forward only)
    # adjoint pass beyond 256 points per path (the reference's GPU limit is (len-1) * 2^d < 1024)
    ("gram_bwd_len300_d1", "gram_bwd", ("rbf", 1.0), 1, False, "bm", 50, (2, 300, 2), (2, 40, 2), {}),
    ("gram_bwd_len600_d0", "gram_bwd", ("rbf", 1.0), 0, False, "bm", 51, (2, 600, 3), (2, 33, 3), {}),
    ("gram_bwd_len1000_d0", "gram_bwd", ("linear", 1.0), 0, False, "bm", 52, (1, 1000, 2), (2, 25, 2), {}),
    ("kernel_bwd_rbf_d0", "kernel_bwd", ("rbf", 0.7), 0, False, "rand", 53, (4, 12, 3), (4, 9, 3), {}),
]


def run_case(case):
    name, op, kspec, d, naive, kind, seed, sx, sy, extra = case
    X, Y = gen(kind, seed, sx, sy)
    if "sub" in extra:
        X, Y = X[:extra["sub"]].clone(), Y[:extra["sub"]].clone()
    if op in ("gram_sym", "gram_sym_bwd"):
        Y = X.clone()
    sk = ref.SigKernel(make_kernel(kspec), d, _naive_solver=naive)
    out = {}
    if op == "kernel":
        out["K"] = sk.compute_kernel(X, Y).numpy()
    elif op == "gram":
        out["G"] = sk.compute_Gram(X, Y, sym=False).numpy()
    elif op == "gram_sym":
        out["G"] = sk.compute_Gram(X, X, sym=True).numpy()
    elif op == "mmd":
        out["mmd"] = sk.compute_mmd(X, Y).numpy()
    elif op == "distance":
        out["dist"] = sk.compute_distance(X, Y).numpy()
    elif op == "scoring":
        out["score"] = sk.compute_scoring_rule(X, Y).numpy()
    elif op == "kernel_bwd":
        Xg = X.clone().requires_grad_(True)
        K = sk.compute_kernel(Xg, Y)
        w = torch.linspace(0.5, 1.5, K.numel(), dtype=torch.float64).reshape(K.shape)
        (K * w).sum().backward()
        out["K"], out["w"], out["grad"] = K.detach().numpy(), w.numpy(), Xg.grad.numpy()
    elif op in ("gram_bwd", "gram_sym_bwd"):
        Xg = X.clone().requires_grad_(True)
        sym = op == "gram_sym_bwd"
        G = sk.compute_Gram(Xg, Xg if sym else Y, sym=sym)
        w = torch.linspace(0.5, 1.5, G.numel(), dtype=torch.float64).reshape(G.shape)
        if sym:
            w = 0.5 * (w + w.T)   # the reference doubles the X-gradient assuming a symmetric upstream grad
        (G * w).sum().backward()
        out["G"], out["w"], out["grad"] = G.detach().numpy(), w.numpy(), Xg.grad.numpy()
    elif op == "mmd_bwd":
        Xg = X.clone().requires_grad_(True)
        m = sk.compute_mmd(Xg, Y)
        m.backward()
        out["mmd"], out["grad"] = m.detach().numpy(), Xg.grad.numpy()
    else:
        raise ValueError(op)
    meta = dict(name=name, op=op, static=kspec[0], param=(kspec[1] if len(kspec) > 1 else 1.0),
                params=list(kspec[1:]), dyadic_order=d, naive=naive, data=kind, seed=seed)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), X=X.numpy(), Y=Y.numpy(),
                        meta=json.dumps(meta), **out)
    return meta


# Directional derivatives (sigkernel.py:43-89, 504-593).  The reference's CPU branch of k_kgrad is broken
# (sigkernel.py:588 unpacks three arrays from a function returning one, cython_backend.pyx:176) and its GPU
# branch needs Numba-CUDA on a device, so these fixtures run the reference's OWN kernel
# `sigkernel_derivatives_Gram_cuda` (cuda_backend.py:165-223) under Numba's CUDA simulator, fed by the
# reference's own static kernels and `tile`, with k_kgrad's increment build (sigkernel.py:524-544) applied
# line by line.  M_inc is zero-padded by one row / column: the reference launches len+1 threads that read
# one element out of bounds.  Run with NUMBA_ENABLE_CUDASIM=1:
#     NUMBA_ENABLE_CUDASIM=1 python tests/golden/make_golden.py deriv
DERIV_CASES = [
    ("deriv_rbf_d1", ("rbf", 0.5), 1, "rand", 30, (2, 5, 2), (3, 4, 2)),
    ("deriv_rbf_d0", ("rbf", 1.0), 0, "randn", 31, (3, 6, 3), (2, 7, 3)),
    ("deriv_linear_d2", ("linear", 1.0), 2, "bm", 32, (2, 4, 2), (2, 5, 2)),
]


def run_deriv_case(case):
    from sigkernel.cuda_backend import sigkernel_derivatives_Gram_cuda
    from sigkernel.sigkernel import tile
    name, kspec, d, kind, seed, sx, sy = case
    X, Y = gen(kind, seed, sx, sy)
    g = torch.Generator().manual_seed(seed + 1000)
    gamma = torch.rand(sx, dtype=torch.float64, generator=g)
    sk = make_kernel(kspec)
    eps = 1e-4
    A, M, _ = sx
    B, N, _ = sy
    MM, NN = (2 ** d) * (M - 1), (2 ** d) * (N - 1)

    def d2(G):
        return G[:, :, 1:, 1:] + G[:, :, :-1, :-1] - G[:, :, 1:, :-1] - G[:, :, :-1, 1:]

    G = sk.Gram_matrix(X, Y)
    G_ = d2(G)
    Gd1 = -(1. / eps) * G
    Gd2 = (1. / eps) * sk.Gram_matrix(X + eps * gamma, Y)
    Gd_ = d2(Gd1) + d2(Gd2)
    Gdd1 = -(1. / eps) * Gd1
    Gdd2 = -(2. / eps) * Gd2
    Gdd3 = (1. / eps ** 2) * sk.Gram_matrix(X + 2. * eps * gamma, Y)
    Gdd_ = d2(Gdd1) + d2(Gdd2) + d2(Gdd3)

    def refine(T):
        return tile(tile(T, 2, 2 ** d) / float(2 ** d), 3, 2 ** d) / float(2 ** d)

    def pad(T):
        return np.pad(T.numpy(), ((0, 0), (0, 0), (0, 1), (0, 1)))

    K = np.zeros((A, B, MM + 2, NN + 2))
    Kd, Kdd = np.zeros_like(K), np.zeros_like(K)
    K[:, :, 0, :] = 1.
    K[:, :, :, 0] = 1.
    tpb = max(MM + 1, NN + 1)
    sigkernel_derivatives_Gram_cuda[(A, B), tpb](pad(refine(G_)), pad(refine(Gd_)), pad(refine(Gdd_)),
                                                 MM + 1, NN + 1, 2 * tpb - 1, K, Kd, Kdd)
    meta = dict(name=name, op="deriv", static=kspec[0], param=kspec[1], dyadic_order=d, data=kind, seed=seed, eps=eps)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), X=X.numpy(), Y=Y.numpy(), gamma=gamma.numpy(),
                        meta=json.dumps(meta), K=K[:, :, MM, NN], K_diff=Kd[:, :, MM, NN], K_diffdiff=Kdd[:, :, MM, NN])
    return meta


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "deriv":
        assert os.environ.get("NUMBA_ENABLE_CUDASIM") == "1", "set NUMBA_ENABLE_CUDASIM=1"
        for c in DERIV_CASES:
            m = run_deriv_case(c)
            print("wrote", m["name"], m["op"])
        sys.exit(0)
    only = sys.argv[1:]
    for c in CASES:
        if only and not any(c[0].startswith(o) for o in only):
            continue
        m = run_case(c)
        print("wrote", m["name"], m["op"])
