"""Quick forward timing sweep on the GPU box (development aid): python tools/time_fwd.py [cfg]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb  # noqa: E402

CFG = {"cfg2": (64, 64, 32, 3, 1), "cfg3": (128, 128, 64, 5, 2), "cfg4f": (128, 128, 64, 3, 1),
       "cfg5s": (64, 512, 128, 8, 2)}


def time_it(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    names = sys.argv[1:] or list(CFG)
    for name in names:
        A, B, L, D, d = CFG[name]
        g = torch.Generator().manual_seed(0)
        X = torch.rand((A, L, D), dtype=torch.float64, generator=g).cuda()
        Y = torch.rand((B, L, D), dtype=torch.float64, generator=g).cuda()
        for kind, par in (("rbf", 0.5), ("linear", 1.0)):
            for w in [int(x) for x in os.environ.get("SKB_WPSM", "0,8,12,16").split(",")]:
                skb._lib.lib.skb_set_warps_per_sm(w)
                best, med = time_it(lambda: skb.ops.sigkernel_forward(X, Y, kind, par, d, "gram"))
                cells = A * B * ((L - 1) << d) ** 2
                print(f"{name} {kind} wpsm={w}: best {best:.3f} ms  med {med:.3f} ms  "
                      f"{A*B/best*1e3:.3e} pairs/s  {cells/best*1e3/1e9:.1f} Gcell/s", flush=True)
        skb._lib.lib.skb_set_warps_per_sm(0)


if __name__ == "__main__":
    main()
