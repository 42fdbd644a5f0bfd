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
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "_ref"))
sys.path.insert(1, ROOT)

CFG = {"cfg2": (64, 64, 32, 3, 1), "cfg3": (128, 128, 64, 5, 2)}


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_numba_b200.json")
    res = {"what": "unmodified reference, SigKernel(RBFKernel(0.5), d).compute_Gram(X.cuda(), Y.cuda()), fp64, "
                   "torch.rand inputs (seed 0), CUDA events, first call (Numba JIT) excluded"}
    try:
        import numba
        import torch
        res["numba"] = numba.__version__
        res["gpu"] = torch.cuda.get_device_name(0)
        from numba import cuda
        res["numba_cc"] = list(cuda.get_current_device().compute_capability)
        import sigkernel
        res["reference_file"] = sigkernel.__file__
        for name, (A, B, L, D, d) in CFG.items():
            torch.manual_seed(0)
            X = torch.rand((A, L, D), dtype=torch.float64).cuda()
            Y = torch.rand((B, L, D), dtype=torch.float64).cuda()
            sk = sigkernel.SigKernel(sigkernel.RBFKernel(sigma=0.5), d)
            entry = {}
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
            try:
                import sigkernel_b200 as skb
                mine = skb.SigKernel(skb.RBFKernel(0.5), d).compute_Gram(X, Y)
                err = ((mine - G).abs() / (G.abs() + 1)).max().item()
                entry["max_mixed_err_vs_sigkernel_b200"] = err
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
            except Exception as exc:  # noqa: BLE001
                entry["sigkernel_b200_error"] = repr(exc)
            res[name] = entry
        res["status"] = "ran"
    except Exception as exc:  # noqa: BLE001
        res["status"] = "failed"
        res["error"] = repr(exc)
        res["traceback"] = traceback.format_exc()[-4000:]
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res)[:3000])


if __name__ == "__main__":
    main()
