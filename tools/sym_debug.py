"""Debug aid: unordered-pair sweep vs the oracle for several dims."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb
from oracle import sigkernel_oracle as O
from tests._util import make_paths, grad_err
for (A, M, D, d, static) in [(4, 30, 8, 1, "linear"), (4, 30, 8, 1, "rbf"), (4, 30, 5, 1, "linear"), (4, 30, 7, 1, "rbf"), (3, 12, 8, 0, "linear"), (3, 12, 9, 2, "rbf")]:
    X = make_paths("bm", 900 + M, (A, M, D))
    par = 0.8 if static == "rbf" else 1.0
    if not skb.ops.adjoint_sym_supported(M, D, d, static):
        print((A, M, D, d, static), "unsupported"); continue
    Xc = X.cuda()
    G, bctx = skb.ops.sigkernel_forward_ctx(Xc, Xc, static, par, d, "sym")
    ok = O.RBFKernel(par) if static == "rbf" else O.LinearKernel()
    _, gp_ref, _ = O.gram_grad_points_analytic(X, X, ok, d)
    for (wd, wo) in ((0.0, 1.0), (1.0, 0.0)):
        g = skb.ops.sigkernel_backward_vjp(Xc, Xc, static, par, d, "sym", bctx, "sym", w_diag=wd, w_off=wo).cpu()
        coef = torch.full((A, A), wo, dtype=torch.float64) + (wd - wo) * torch.eye(A, dtype=torch.float64)
        expect = 2.0 * torch.einsum('ab,abmd->amd', coef, gp_ref)
        # x-side only reference (sum_b coef d1k) to see which half is off
        half = torch.einsum('ab,abmd->amd', coef, gp_ref)
        err = (g - expect).abs() / (expect.abs() + 1)
        print((A, M, D, d, static), (wd, wo), "max err", float(err.max()), "err per dim", [float(err[..., k].max()) for k in range(D)])
