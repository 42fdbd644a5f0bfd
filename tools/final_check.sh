python - <<'PY'
import torch, sigkernel_b200 as sigkernel
X = torch.rand(128, 64, 3, dtype=torch.float64, device="cuda", requires_grad=True)
Y = torch.rand(128, 64, 3, dtype=torch.float64, device="cuda")
sk = sigkernel.SigKernel(sigkernel.RBFKernel(sigma=0.5), dyadic_order=1)
G = sk.compute_Gram(X, Y)
sk.compute_mmd(X, Y).backward()
k, k_g, k_gg = sk.compute_kernel_and_derivatives_Gram(X.detach(), Y, torch.rand_like(X.detach()))
print("usage ok", tuple(G.shape), tuple(X.grad.shape), tuple(k_gg.shape), bool(torch.isfinite(X.grad).all()))
PY
(timeout 900 python bench.py 2>&1 | tail -1) > gpurun_out/r02_bench_n1.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read())
print("N=1", d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['cfg4']['ms_per_step'], d['cfg4']['gram_with_grad_points_ms'], d['gpu_launches'], d['clocks'])
PY
