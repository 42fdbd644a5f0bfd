"""Measure the fp64 issue rate of the GPU (the solver's roofline denominator) with skb_fp64_probe.
Writes gpurun_out/fp64_peak.json.  Run on the GPU box: python tools/probe_fp64.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb  # noqa: E402


def measure(op, blocks, threads, iters, reps=5):
    lib = skb._lib.lib
    sink = torch.zeros(8, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    best = 1e30
    for _ in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        skb._lib.check(lib.skb_fp64_probe(op, blocks, threads, iters, sink.data_ptr(), st))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    n = blocks * threads * iters * 16
    return n / (best * 1e-3), best


def main():
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    out = {"gpu": torch.cuda.get_device_name(0), "sms": sms, "ops": {}}
    for name, op in (("dfma", 0), ("dadd", 1), ("dmul", 2), ("dfma_3distinct", 3), ("dfma_2distinct", 4), ("dadd_2distinct", 5), ("dfma_a_v_z", 6), ("dadd+lop3", 7), ("dadd+2lop3", 8), ("dadd+imad_imm", 9)):
        for wpsm in (16, 64):
            blocks = sms * (wpsm * 32 // 256) if wpsm * 32 >= 256 else sms
            threads = 256 if wpsm * 32 >= 256 else wpsm * 32
            rate, ms = measure(op, blocks, threads, 200000)
            out["ops"].setdefault(name, {})[f"warps_per_sm_{wpsm}"] = {"dp_instr_per_s": rate, "ms": ms}
            print(name, wpsm, f"{rate/1e12:.3f} T thread-level DP instr/s  ({ms:.2f} ms)", flush=True)
    out["peak_dp_instr_per_s"] = max(v["dp_instr_per_s"] for o in out["ops"].values() for v in o.values())
    out["nominal_dp_instr_per_s_at_1965MHz"] = sms * 64 * 1.965e9
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/fp64_peak.json", "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "ops"}))


if __name__ == "__main__":
    main()
