"""Development aid: run skb_sigkernel_fwd_bwd at a config through the raw ABI and report the reconstruction flag."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb  # noqa: E402

CFG = {"cfg4": (128, 128, 64, 3, 1), "cfg3b": (128, 128, 64, 5, 2), "small": (6, 5, 40, 3, 1)}
lib, chk = skb._lib.lib, skb._lib.check
if os.environ.get("SKB_ADJ_MODE"):
    lib.skb_set_adjoint_mode(int(os.environ["SKB_ADJ_MODE"]))
for name in (sys.argv[1:] or ["cfg4"]):
    A, B, L, D, d = CFG[name]
    g = torch.Generator().manual_seed(0)
    X = torch.rand((A, L, D), dtype=torch.float64, generator=g).cuda()
    Y = torch.rand((B, L, D), dtype=torch.float64, generator=g).cuda()
    n = lib.skb_bwd_workspace_bytes(A, B, L, L, D, d, 0)
    ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
    out = torch.empty(A * B, dtype=torch.float64, device="cuda")
    gp = torch.empty((A * B, L, D), dtype=torch.float64, device="cuda")
    for rep in range(3):
        chk(lib.skb_sigkernel_fwd_bwd(X.data_ptr(), Y.data_ptr(), 0, A, B, L, L, D, d, 1, 0.5, 0, 0, out.data_ptr(), gp.data_ptr(),
                                      ws.data_ptr(), n, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    print(name, "workspace", n, "flag", int(ws[64:68].view(torch.int32).item()), "plan", lib.skb_adjoint_plan(L, L, D, d, 1, 0), flush=True)
