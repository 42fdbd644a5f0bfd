python - <<'PY'
import torch, sigkernel_b200 as skb, sys
sys.path.insert(0,'.')
from tools.time_fwd import time_it
g = torch.Generator().manual_seed(0)
X = torch.rand((128, 64, 3), dtype=torch.float64, generator=g).cuda()
Y = torch.rand((128, 64, 3), dtype=torch.float64, generator=g).cuda()
sk = skb.SigKernel(skb.RBFKernel(0.5), 1)
for mode in (-1, 3, -1, 3):
    skb._lib.lib.skb_set_adjoint_mode(mode)
    def mmd():
        Xg = X.clone().requires_grad_(True)
        sk.compute_mmd(Xg, Y).backward()
        return Xg.grad
    b, m = time_it(mmd, reps=10)
    res = skb.ops.sigkernel_forward_ctx(X, X, "rbf", 0.5, 1, "sym")
    if mode == -1:
        t, _ = time_it(lambda: skb.ops.sigkernel_backward_vjp(X, X, "rbf", 0.5, 1, "sym", res[1], "sym", w_diag=0.0, w_off=1e-4))
    else:
        r2 = skb.ops.sigkernel_forward_ctx(X, X, "rbf", 0.5, 1, "gram")
        t, _ = time_it(lambda: skb.ops.sigkernel_backward_vjp(X, X, "rbf", 0.5, 1, "gram", r2[1], "gram", w_diag=0.0, w_off=1e-4, out_scale=2.0))
    print(f"adjoint mode {mode}: compute_mmd+backward best {b:.3f} ms med {m:.3f}; K_XX backward alone {t:.3f} ms", flush=True)
skb._lib.lib.skb_set_adjoint_mode(-1)
PY
