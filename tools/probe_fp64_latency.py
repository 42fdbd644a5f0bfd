"""Dependent-issue latency of DFMA on this GPU (development aid): cycles per instruction with C independent chains
per thread and W warps per scheduler.  python tools/probe_fp64_latency.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb  # noqa: E402
from tools.probe_fp64 import measure  # noqa: E402


def main():
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    clk = 1.92e9
    for wps in (1, 2, 4):
        for name, op, c in (("1 chain", 10, 1), ("2 chains", 11, 2), ("4 chains", 12, 4), ("8 chains", 13, 8), ("16 chains", 0, 16)):
            rate, ms = measure(op, sms, 128 * wps, 100000)
            # warp instructions per scheduler per second = rate / 32 / (sms * 4)
            cyc = clk / (rate / 32 / (sms * 4))
            print(f"warps/scheduler {wps} {name}: {cyc:.2f} cycles per warp instruction per scheduler (at {clk/1e9} GHz), per chain step {cyc * c * wps:.1f}", flush=True)


if __name__ == "__main__":
    main()
