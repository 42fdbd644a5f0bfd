import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb
from tests._util import make_paths
lib, chk = skb._lib.lib, skb._lib.check
torch.set_printoptions(precision=5, linewidth=220)


def raw(X, Y, static, par, d, mode):
    lib.skb_set_adjoint_mode(mode)
    A, M, D = X.shape
    B, N, _ = Y.shape
    n = lib.skb_bwd_workspace_bytes(A, B, M, N, D, d, 0)
    ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
    out = torch.empty(A * B, dtype=torch.float64, device="cuda")
    gp = torch.empty((A * B, M, D), dtype=torch.float64, device="cuda")
    chk(lib.skb_sigkernel_fwd_bwd(X.data_ptr(), Y.data_ptr(), 0, A, B, M, N, D, d, 1 if static == "rbf" else 0, par, 0, 0,
                                  out.data_ptr(), gp.data_ptr(), ws.data_ptr(), n, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    lib.skb_set_adjoint_mode(-1)
    flag = int(ws[64:68].view(torch.int32).item())
    # boundary context sits after the prepared paths
    Dp = 4 if D + 1 <= 4 else (6 if D + 1 <= 6 else 10)
    xb = (A * M * Dp * 8 + 255) // 256 * 256
    yb = (B * N * Dp * 8 + 255) // 256 * 256
    off = 256 + 2 * xb + 2 * yb
    MM, NN = (M - 1) << d, (N - 1) << d
    bs, cs = (NN + 1 + 3) // 4 * 4, (MM + 1 + 3) // 4 * 4
    ctx = ws[off:off + (bs + cs) * 8 * A * B].view(torch.float64)
    brow = ctx[:bs][:NN + 1].clone()
    bcol = ctx[A * B * bs:A * B * bs + cs][:MM + 1].clone()
    return out, gp, flag, brow, bcol


for (M, N, D, d, static) in [(9, 7, 3, 1, "linear"), (40, 7, 2, 0, "linear"), (5, 6, 2, 0, "linear")]:
    X = make_paths("rand", 1, (1, M, D)).cuda(); Y = make_paths("rand", 2, (1, N, D)).cuda()
    par = 0.7 if static == "rbf" else 1.0
    o0, g0, f0, _, _ = raw(X, Y, static, par, d, 0)
    o1, g1, f1, brow, bcol = raw(X, Y, static, par, d, 1)
    print("====", (M, N, D, d, static), "flags", f0, f1, "G", o0.item(), o1.item())
    print("brow", brow)
    print("bcol", bcol)
    print("stored", g0[0, :, 0])
    print("recon ", g1[0, :, 0])
