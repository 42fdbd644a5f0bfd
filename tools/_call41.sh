timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python - <<'PY'
import torch, sigkernel_b200 as skb, sys
sys.path.insert(0,'.')
from tools.time_fwd import time_it
g = torch.Generator().manual_seed(0)
X = torch.rand((128, 64, 5), dtype=torch.float64, generator=g).cuda()
Y = torch.rand((128, 64, 5), dtype=torch.float64, generator=g).cuda()
for naive in (False, True):
    b, m = time_it(lambda: skb.ops.sigkernel_forward(X, Y, "rbf", 0.5, 2, "gram", naive))
    print(f"cfg3 forward naive={naive}: best {b:.3f} ms", flush=True)
PY
