"""K_XX backward through the unordered-pair sweep at the cfg4 shape (profiling aid): python tools/run_sym.py [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb
g = torch.Generator().manual_seed(0)
X = torch.rand((128, 64, 3), dtype=torch.float64, generator=g).cuda()
res = skb.ops.sigkernel_forward_ctx(X, X, "rbf", 0.5, 1, "sym")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    gx = skb.ops.sigkernel_backward_vjp(X, X, "rbf", 0.5, 1, "sym", res[1], "sym", w_diag=0.0, w_off=1e-4)
torch.cuda.synchronize()
print("done", float(gx.abs().sum()))
