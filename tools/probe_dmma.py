"""Does mma.sync m8n8k4 f64 (DMMA) run beside the DFMA pipe on this GPU?  (development aid)  python tools/probe_dmma.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sigkernel_b200 as skb  # noqa: E402
from tools.probe_fp64 import measure  # noqa: E402


def main():
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for wps in (2, 4):
        res = {}
        for name, op in (("16 DMMA / iter", 20), ("16 DFMA + 4 DMMA / iter", 21), ("16 DFMA / iter", 22)):
            rate, ms = measure(op, sms * wps // 2, 256, 100000)
            res[name] = ms
            print(f"warps/scheduler {wps}: {name}: {ms:.3f} ms", flush=True)
        t_dfma, t_mix, t_dmma = res["16 DFMA / iter"], res["16 DFMA + 4 DMMA / iter"], res["16 DMMA / iter"]
        print(f"  one DMMA costs {t_dmma / t_dfma:.2f} DFMA issue slots when alone; mixed / (DFMA + DMMA/4 alone) = {t_mix / (t_dfma + t_dmma / 4):.2f} "
              f"(1.0 = same pipe, {max(t_dfma, t_dmma / 4) / (t_dfma + t_dmma / 4):.2f} = fully overlapped)", flush=True)


if __name__ == "__main__":
    main()
