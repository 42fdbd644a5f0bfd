"""Launch one BASELINE config's forward a few times (target for ncu).  python tools/run_cfg.py cfg3 [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb  # noqa: E402

CFG = {"cfg2": (64, 64, 32, 3, 1), "cfg3": (128, 128, 64, 5, 2), "cfg4f": (128, 128, 64, 3, 1),
       "cfg5s": (64, 512, 128, 8, 2)}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "fwd"
if os.environ.get('SKB_TILE_MODE'):
    skb._lib.lib.skb_set_tile_mode(int(os.environ['SKB_TILE_MODE']))
A, B, L, D, d = CFG[name]
g = torch.Generator().manual_seed(0)
X = torch.rand((A, L, D), dtype=torch.float64, generator=g).cuda()
Y = torch.rand((B, L, D), dtype=torch.float64, generator=g).cuda()
for _ in range(reps):
    if mode == "fwd":
        out = skb.ops.sigkernel_forward(X, Y, "rbf", 0.5, d, "gram")
    else:
        out = skb.ops.sigkernel_forward_backward(X, Y, "rbf", 0.5, d, "gram")
torch.cuda.synchronize()
print("done", float(out[0].sum() if isinstance(out, tuple) else out.sum()))
