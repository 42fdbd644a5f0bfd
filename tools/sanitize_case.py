"""Small forward / backward / derivative calls for compute-sanitizer (memcheck, racecheck)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb

g = torch.Generator().manual_seed(0)
for (A, B, M, N, D, d) in [(5, 4, 9, 4, 2, 1), (2, 2, 100, 11, 8, 1), (1, 2, 250, 6, 3, 2)]:
    X = torch.rand((A, M, D), dtype=torch.float64, generator=g).cuda()
    Y = torch.rand((B, N, D), dtype=torch.float64, generator=g).cuda()
    for k in (skb.RBFKernel(0.5), skb.LinearKernel()):
        G = skb.SigKernel(k, d).compute_Gram(X, Y)
X = torch.rand((3, 12, 3), dtype=torch.float64, generator=g).cuda().requires_grad_(True)
Y = torch.rand((4, 10, 3), dtype=torch.float64, generator=g).cuda()
sk = skb.SigKernel(skb.RBFKernel(0.5), 1)
sk.compute_mmd(X, Y).backward()
gam = torch.rand((3, 12, 3), dtype=torch.float64, generator=g).cuda()
sk.compute_kernel_and_derivatives_Gram(X.detach(), Y, gam)
# round 2: unordered-pair sweep (mmd above used it for K_XX), scoring rule, eager backward of a long path (materialised
# grids), reconstruction with 2 warps per pair, streaming derivative kernel at 8 rows per lane, symmetric forward slice
Xs = torch.rand((5, 40, 3), dtype=torch.float64, generator=g).cuda().requires_grad_(True)
sk.compute_scoring_rule(Xs, Y[0:1]).backward()
Xl = torch.rand((2, 300, 2), dtype=torch.float64, generator=g).cuda()
skb.ops.sigkernel_forward_backward(Xl, Y[:2, :, :2].contiguous(), "rbf", 0.5, 2, "gram")
Xm = torch.rand((2, 150, 3), dtype=torch.float64, generator=g).cuda().requires_grad_(True)
(skb.SigKernel(skb.RBFKernel(0.5), 1).compute_Gram(Xm, Y) ** 2).sum().backward()
gam2 = torch.rand((2, 60, 3), dtype=torch.float64, generator=g).cuda()
skb.SigKernel(skb.RBFKernel(0.5), 2).compute_kernel_and_derivatives_Gram(Xm.detach()[:, :60].contiguous(), Y, gam2)
out = torch.zeros((5, 5), dtype=torch.float64, device="cuda")
skb.ops.sigkernel_forward_range(Xs.detach(), Xs.detach(), "rbf", 0.5, 1, 3, 11, "sym", out=out)
torch.cuda.synchronize()
print("ok", float(G.sum()), float(X.grad.abs().sum()), float(Xs.grad.abs().sum()), float(Xm.grad.abs().sum()), float(out.sum()))
