"""Tile kernel vs fwd5_kernel on the GPU box (development aid): python tools/time_tile.py [cfg ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb  # noqa: E402
from tools.time_fwd import CFG, time_it  # noqa: E402


def main():
    lib = skb._lib.lib
    for name in (sys.argv[1:] or ["cfg3", "cfg5s", "cfg4f"]):
        A, B, L, D, d = CFG[name]
        g = torch.Generator().manual_seed(0)
        X = torch.rand((A, L, D), dtype=torch.float64, generator=g).cuda()
        Y = torch.rand((B, L, D), dtype=torch.float64, generator=g).cuda()
        res = {}
        for mode in (0, 1):
            lib.skb_set_tile_mode(mode)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(); ev1.record(); torch.cuda.synchronize()
            best, med = time_it(lambda: skb.ops.sigkernel_forward(X, Y, "rbf", 0.5, d, "gram"), reps=20)
            lib.skb_set_profile_events(ev0.cuda_event, ev1.cuda_event)
            ks = []
            for _ in range(10):
                res[mode] = skb.ops.sigkernel_forward(X, Y, "rbf", 0.5, d, "gram")
                torch.cuda.synchronize()
                ks.append(ev0.elapsed_time(ev1))
            lib.skb_set_profile_events(None, None)
            MM = (L - 1) << d
            frac4 = 4.0 * A * B * MM * MM / (min(ks) * 1e-3) / 1.84e13
            print(f"{name} tile_mode={mode}: op best {best:.4f} ms med {med:.4f} ms; solver kernel best {min(ks):.4f} ms "
                  f"(4/cell fraction of 1.84e13: {frac4:.3f})", flush=True)
        err = ((res[0] - res[1]).abs() / (res[0].abs() + 1)).max().item()
        print(f"{name}: max mixed error tile vs fwd5 {err:.2e}", flush=True)
    lib.skb_set_tile_mode(-1)


if __name__ == "__main__":
    main()
