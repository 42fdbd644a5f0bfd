"""Quick backward timing on the GPU box (development aid): python tools/time_bwd.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb  # noqa: E402
from tools.time_fwd import time_it  # noqa: E402

CFG = {"cfg4": (128, 128, 64, 3, 1), "cfg3b": (128, 128, 64, 5, 2), "cfg2b": (64, 64, 32, 3, 1)}


def main():
    if os.environ.get('SKB_ADJ_MODE'):
        skb._lib.lib.skb_set_adjoint_mode(int(os.environ['SKB_ADJ_MODE']))
    for name in (sys.argv[1:] or list(CFG)):
        A, B, L, D, d = CFG[name]
        g = torch.Generator().manual_seed(0)
        X = torch.rand((A, L, D), dtype=torch.float64, generator=g).cuda()
        Y = torch.rand((B, L, D), dtype=torch.float64, generator=g).cuda()
        best, med = time_it(lambda: skb.ops.sigkernel_forward_backward(X, Y, "rbf", 0.5, d, "gram"))
        print(f"{name} fwd+bwd gram: best {best:.3f} ms med {med:.3f} ms  {A*B/best*1e3:.3e} pairs/s", flush=True)
        bf, mf = time_it(lambda: skb.ops.sigkernel_forward(X, Y, "rbf", 0.5, d, "gram"))
        print(f"{name} fwd only: best {bf:.3f} ms", flush=True)
        if name == "cfg4":
            sk = skb.SigKernel(skb.RBFKernel(0.5), d)

            def mmd():
                Xg = X.clone().requires_grad_(True)
                sk.compute_mmd(Xg, Y).backward()
                return Xg.grad
            best, med = time_it(mmd, reps=5)
            print(f"{name} compute_mmd+backward (public API): best {best:.3f} ms med {med:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
