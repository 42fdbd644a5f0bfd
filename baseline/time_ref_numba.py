"""Time the UNMODIFIED reference (crispitagorico/sigkernel installed under baseline/_ref by
`pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`)
through its own public API on CUDA tensors on this GPU: `SigKernel.compute_Gram` -> `_SigKernelGram.forward`
-> Numba `sigkernel_Gram_cuda` (sigkernel/cuda_backend.py:121-160, sigkernel/sigkernel.py:366-382).

    python baseline/time_ref_numba.py [out.json]

numba 0.65 lists compute capabilities up to 9.0 only, so on a B200 (sm_100) the kernel can only run through the
driver's PTX JIT; whatever happens (times or the exception) is written to the JSON file.  Not product, not oracle:
the "reference's own cuda_backend" denominator of BASELINE.json's north_star."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "_ref"))
sys.path.insert(1, ROOT)

CFG = {"cfg2": (64, 64, 32, 3, 1), "cfg3": (128, 128, 64, 5, 2)}


def one_config(name, warm_pool):
    """Run one config in THIS process and return its entry (called in a subprocess by main(): the reference's kernel
    reads its increment tensor one element out of bounds per row, which at the headline config runs off
    the end of a 2.1 GB allocation into unmapped memory -- cudaErrorIllegalAddress kills the context)."""
    import numba
    import torch
    sys.path.insert(0, os.path.join(HERE, "_ref"))
    import sigkernel
    A, B, L, D, d = CFG[name]
    if warm_pool:
        # environment workaround, the reference stays unmodified: let torch's caching allocator own one large block first,
        # so that the tensors of the run are carved out of it and the stray read lands in mapped memory
        pool = torch.empty(int(warm_pool) << 30, dtype=torch.uint8, device="cuda")
        del pool
    torch.manual_seed(0)
    X = torch.rand((A, L, D), dtype=torch.float64).cuda()
    Y = torch.rand((B, L, D), dtype=torch.float64).cuda()
    sk = sigkernel.SigKernel(sigkernel.RBFKernel(sigma=0.5), d)
    entry = {"warm_pool_GiB": warm_pool}
    for label, mb in (("max_batch_100", 100), ("max_batch_full", max(A, B))):
        t0 = time.perf_counter()
        G = sk.compute_Gram(X, Y, sym=False, max_batch=mb)
        torch.cuda.synchronize()
        entry[label + "_first_call_s"] = time.perf_counter() - t0
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            G = sk.compute_Gram(X, Y, sym=False, max_batch=mb)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        entry[label + "_ms_best"] = ts[0]
        entry[label + "_ms_median"] = ts[len(ts) // 2]
        entry[label + "_pairs_per_s"] = A * B / (ts[0] * 1e-3)
    import sigkernel_b200 as skb
    mine = skb.SigKernel(skb.RBFKernel(0.5), d).compute_Gram(X, Y)
    entry["max_mixed_err_vs_sigkernel_b200"] = ((mine - G).abs() / (G.abs() + 1)).max().item()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        mine = skb.SigKernel(skb.RBFKernel(0.5), d).compute_Gram(X, Y)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    entry["sigkernel_b200_ms_best"] = min(ts)
    return entry


def main():
    if len(sys.argv) > 3 and sys.argv[1] == "--one":
        try:
            print("ENTRY " + json.dumps(one_config(sys.argv[2], int(sys.argv[3]))))
        except Exception as exc:  # noqa: BLE001
            print("ENTRY " + json.dumps({"warm_pool_GiB": int(sys.argv[3]), "error": repr(exc)[:600]}))
        return
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_numba_b200.json")
    import subprocess
    res = {"what": "unmodified reference, SigKernel(RBFKernel(0.5), d).compute_Gram(X.cuda(), Y.cuda()), fp64, "
                   "torch.rand inputs (seed 0), CUDA events, first call (Numba JIT) excluded; one subprocess per config and "
                   "attempt (plain first, then with a pre-warmed allocator pool)"}
    for name in CFG:
        attempts = []
        for pool in (0, 40):
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name, str(pool)], capture_output=True, text=True)
            ent = None
            for ln in out.stdout.splitlines():
                if ln.startswith("ENTRY "):
                    ent = json.loads(ln[6:])
            if ent is None:
                ent = {"warm_pool_GiB": pool, "error": "process died: " + (out.stderr or "")[-400:]}
            attempts.append(ent)
            if "error" not in ent:
                break
        res[name] = attempts
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res)[:3000])
    return


if __name__ == "__main__":
    main()
