"""Timing of compute_kernel_and_derivatives_Gram (development aid): python tools/time_deriv.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb  # noqa: E402
from tools.time_fwd import time_it  # noqa: E402

for (A, B, L, D, d) in [(64, 64, 32, 3, 1), (128, 128, 64, 5, 2)]:
    g = torch.Generator().manual_seed(0)
    X = torch.rand((A, L, D), dtype=torch.float64, generator=g).cuda()
    Y = torch.rand((B, L, D), dtype=torch.float64, generator=g).cuda()
    gam = torch.rand((A, L, D), dtype=torch.float64, generator=g).cuda()
    sk = skb.SigKernel(skb.RBFKernel(0.5), d)
    best, med = time_it(lambda: sk.compute_kernel_and_derivatives_Gram(X, Y, gam), reps=5)
    eps = 1e-4
    K0, K1, K2 = (sk.static_kernel.Gram_matrix(X + s * eps * gam, Y) for s in (0., 1., 2.))
    bk, _ = time_it(lambda: skb.ops.kernel_and_derivatives_from_static(K0, K1, K2, d, eps), reps=5)
    print(f"{A}x{B} len {L} dim {D} d {d}: public API best {best:.3f} ms (med {med:.3f}); solver entry alone {bk:.3f} ms "
          f"= {A*B/bk*1e3:.3e} pairs/s", flush=True)
