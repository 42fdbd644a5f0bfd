"""Time the STAND-IN for the reference's GPU path (baseline/ref_cuda_standin.cu) on this GPU, with the torch ops
the reference runs in front of its kernel (static_kernels.py:58-73, sigkernel.py:362-364, 370-382, 607-613) and its
default max_batch=100 splitting (sigkernel.py:102-127), next to sigkernel_b200 on the same inputs.

    python baseline/time_ref_standin.py [cfg2|cfg3]

Builds baseline/libref_standin.so with nvcc on first use.  Labelled stand-in, see the .cu header and BASELINE.md."""
import ctypes
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import sigkernel_b200 as skb  # noqa: E402

CFG = {"cfg2": (64, 64, 32, 3, 1), "cfg3": (128, 128, 64, 5, 2)}


def lib():
    so = os.path.join(HERE, "libref_standin.so")
    if not os.path.exists(so):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC",
                               os.path.join(HERE, "ref_cuda_standin.cu"), "-o", so])
    L = ctypes.CDLL(so)
    L.ref_gram_launch.restype = ctypes.c_int
    L.ref_gram_launch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_int, ctypes.c_void_p]
    return L


def tile(x, dim, n):          # what the reference's tile() computes (repeat_interleave), sigkernel.py:607-613
    return torch.repeat_interleave(x, n, dim=dim)


def ref_gram_block(L, X, Y, sigma, d):
    """_SigKernelGram.forward on the CUDA branch (sigkernel.py:349-401) with the stand-in kernel."""
    A, M, D = X.shape
    B, N, _ = Y.shape
    MM, NN = (2 ** d) * (M - 1), (2 ** d) * (N - 1)
    xs = (X ** 2).sum(2)[:, None, :, None]
    ys = (Y ** 2).sum(2)[None, :, None, :]
    G_static = torch.exp(-(-2. * torch.einsum('ipk,jqk->ijpq', X, Y) + xs + ys) / sigma)
    G_ = G_static[:, :, 1:, 1:] + G_static[:, :, :-1, :-1] - G_static[:, :, 1:, :-1] - G_static[:, :, :-1, 1:]
    G_ = tile(tile(G_, 2, 2 ** d) / float(2 ** d), 3, 2 ** d) / float(2 ** d)
    G_ = torch.nn.functional.pad(G_, (0, 1, 0, 1))               # the out-of-bounds row/column the reference reads
    sol = torch.zeros((A, B, MM + 2, NN + 2), device=X.device, dtype=X.dtype)
    sol[:, :, 0, :] = 1.
    sol[:, :, :, 0] = 1.
    rc = L.ref_gram_launch(G_.data_ptr(), A, B, MM + 1, NN + 1, sol.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    return sol[:, :, MM, NN]


def ref_gram(L, X, Y, sigma, d, max_batch=100):
    """SigKernel.compute_Gram's recursive halving (sigkernel.py:102-127)."""
    A, B = X.shape[0], Y.shape[0]
    if A <= max_batch and B <= max_batch:
        return ref_gram_block(L, X, Y, sigma, d)
    if A <= max_batch:
        c = B // 2
        return torch.cat((ref_gram(L, X, Y[:c], sigma, d, max_batch), ref_gram(L, X, Y[c:], sigma, d, max_batch)), 1)
    if B <= max_batch:
        c = A // 2
        return torch.cat((ref_gram(L, X[:c], Y, sigma, d, max_batch), ref_gram(L, X[c:], Y, sigma, d, max_batch)), 0)
    ca, cb = A // 2, B // 2
    top = torch.cat((ref_gram(L, X[:ca], Y[:cb], sigma, d, max_batch), ref_gram(L, X[:ca], Y[cb:], sigma, d, max_batch)), 1)
    bot = torch.cat((ref_gram(L, X[ca:], Y[:cb], sigma, d, max_batch), ref_gram(L, X[ca:], Y[cb:], sigma, d, max_batch)), 1)
    return torch.cat((top, bot), 0)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), out


def main():
    L = lib()
    for name in (sys.argv[1:] or ["cfg2", "cfg3"]):
        A, B, Lp, D, d = CFG[name]
        g = torch.Generator().manual_seed(0)
        X = torch.rand((A, Lp, D), dtype=torch.float64, generator=g).cuda()
        Y = torch.rand((B, Lp, D), dtype=torch.float64, generator=g).cuda()
        t_ref, G_ref = timeit(lambda: ref_gram(L, X, Y, 0.5, d))
        sk = skb.SigKernel(skb.RBFKernel(0.5), d)
        t_our, G = timeit(lambda: sk.compute_Gram(X, Y), reps=20, warm=3)
        err = float(((G - G_ref).abs() / (G_ref.abs() + 1)).max())
        print(f"{name}: reference-structure stand-in {t_ref:.3f} ms ({A*B/t_ref*1e3:.3e} pairs/s)   sigkernel_b200 {t_our:.3f} ms "
              f"({A*B/t_our*1e3:.3e} pairs/s)   ratio {t_ref/t_our:.1f}x   max mixed error {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
