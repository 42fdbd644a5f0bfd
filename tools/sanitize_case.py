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
torch.cuda.synchronize()
print("ok", float(G.sum()), float(X.grad.abs().sum()))
