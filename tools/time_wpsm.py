"""Residency experiment (development aid): fwd5 (16 lanes per pair) at a batch large enough that tails do not matter,
for several resident-warps-per-SM settings.  python tools/time_wpsm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb  # noqa: E402
from tools.time_fwd import time_it  # noqa: E402

lib = skb._lib.lib
lib.skb_set_tile_mode(0)
for (A, B, L, D, d) in [(256, 256, 64, 5, 2), (128, 128, 64, 5, 2)]:
    g = torch.Generator().manual_seed(0)
    X = torch.rand((A, L, D), dtype=torch.float64, generator=g).cuda()
    Y = torch.rand((B, L, D), dtype=torch.float64, generator=g).cuda()
    for w in (6, 8, 9, 10, 11):
        lib.skb_set_warps_per_sm(w)
        best, med = time_it(lambda: skb.ops.sigkernel_forward(X, Y, "rbf", 0.5, d, "gram"), reps=10)
        MM = (L - 1) << d
        print(f"{A}x{B} wpsm={w}: best {best:.4f} ms  ({A*B/best*1e3:.3e} pairs/s; 4/cell frac {4.0*A*B*MM*MM/(best*1e-3)/1.84e13:.3f})", flush=True)
lib.skb_set_warps_per_sm(0)
lib.skb_set_tile_mode(-1)
