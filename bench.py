#!/usr/bin/env python
"""bench.py -- Gram path-pairs/sec (fp64) of the signature-kernel solver.

    python bench.py --gpus N --steps K --warmup W            (ours; N>1 under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path, host cores)

Workload = BASELINE.json configs[2] (the headline): compute_Gram 128x128, len 64, dim 5, dyadic_order 2,
RBFKernel(sigma=0.5), fp64, synthetic torch.rand paths (README recipe of the reference).  A step is one
pass of the hot path over one batch: the whole Gram matrix, through the public API (SigKernel.compute_Gram; at
N>1 sigkernel_b200.distributed.compute_Gram_sharded: weak scaling, every rank owns 128 rows of X, Y is
replicated, G is reassembled on every rank).  Sub-objects of the same JSON line: `cfg4` (configs[3]:
compute_mmd + backward, N=1) and `cfg5_sharded` (configs[4]: 512x512 sharded over the N ranks, N>1).

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(name="cfg3", A=128, B=128, L=64, D=5, d=2, sigma=0.5)
WORKLOAD = ("compute_Gram 128x128 len=64 dim=5 dyadic_order=2 RBFKernel(sigma=0.5) fp64 "
            "(BASELINE.json configs[2], headline)")
METRIC = "gram_path_pairs_per_sec_fp64"
UNIT = "pairs/s"


def make_inputs(rank, torch):
    """README recipe of the reference (README.md:57-59): seeded torch.rand on the CPU."""
    gy = torch.Generator().manual_seed(0)
    gx = torch.Generator().manual_seed(1000 + rank)
    X = torch.rand((CFG["A"], CFG["L"], CFG["D"]), dtype=torch.float64, generator=gx)
    Y = torch.rand((CFG["B"], CFG["L"], CFG["D"]), dtype=torch.float64, generator=gy)
    return X, Y


# DP instructions the algorithm needs per pair in the formulation of skb_fwd5.cuh (DESIGN.md "Roofline"):
#   3 per fine cell (DADD, DMUL, DFMA), 4 per coarse cell (increment 1, -b 1, a 2),
#   per node: D+1 for the dot product (+ norms), 9 for the table-driven exp (reduction 4, polynomial 4,
#   table multiply-add 1) and 1 for the column difference
def dp_instr_per_pair(L, D, d, rbf=True):
    MM = (L - 1) << d
    return 3 * MM * MM + 4 * (L - 1) * (L - 1) + ((D + 1) + 1 + (9 if rbf else 0)) * L * L


def stencil_dp_instr_per_pair(L, d):
    MM = (L - 1) << d
    return 3 * MM * MM


# ----------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi while the timed region runs
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [ln.split(",") for ts, ln in self.lines if t0 <= ts <= t1 + 0.1] or \
               [ln.split(",") for ts, ln in self.lines]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
# the reference arm and the CPU baseline: the reference's own CPU algorithm on the host cores
# ----------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def _have_reference_package():
    return os.path.isdir(os.path.join(REF_DIR, "sigkernel"))


def _cpu_gram_rows(args):
    """One worker: Gram rows [lo, hi) of the workload on the CPU.  With the reference installed under baseline/_ref
    (pip install --target) this is the UNMODIFIED reference through its own public API,
    sigkernel.SigKernel(RBFKernel(0.5), 2).compute_Gram(X[lo:hi], Y); otherwise the oracle port around the reference's
    compiled Cython solver (oracle/_ref) or the C restatement."""
    lo, hi, rank_seed, threads = args
    import torch
    torch.set_num_threads(threads)
    X, Y = make_inputs(rank_seed, torch)
    if _have_reference_package():
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import sigkernel as ref
        sk = ref.SigKernel(ref.RBFKernel(sigma=CFG["sigma"]), CFG["d"])
        return float(sk.compute_Gram(X[lo:hi], Y).sum())
    from oracle import sigkernel_oracle as O
    backend = "ref" if O.ref_backend() is not None else "c"
    out = []
    for r0 in range(lo, hi, 4):                       # 4 rows x 128 columns at a time: ~0.5 GB of grids
        r1 = min(hi, r0 + 4)
        out.append(O.compute_Gram(X[r0:r1], Y, O.RBFKernel(CFG["sigma"]), CFG["d"], backend=backend))
    return torch.cat(out).sum().item()


def cpu_path_description():
    if _have_reference_package():
        return "reference", ("unmodified reference package (baseline/_ref) through its public API "
                             "SigKernel(RBFKernel(0.5), 2).compute_Gram on CPU tensors")
    from oracle import sigkernel_oracle as O
    if O.ref_backend() is not None:
        return "reference-solver+port", "reference Cython solver (oracle/_ref) behind the oracle's torch port of the static kernel and tile()"
    return "port", "oracle port (C restatement of the solver)"


def cpu_gram_throughput(rows, workers, threads=1):
    """pairs/s of the CPU path on `rows` rows of X (x all 128 columns) using `workers` processes of `threads` torch
    threads each (the reference's PDE solve is single-threaded; only its static-kernel ops use torch threads)."""
    import multiprocessing as mp
    kind, _ = cpu_path_description()
    workers = max(1, min(workers, rows))
    bounds = [(rows * w // workers, rows * (w + 1) // workers, 0, threads) for w in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        _cpu_gram_rows(bounds[0])
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            pool.map(_cpu_gram_rows, bounds)
    dt = time.perf_counter() - t0
    return rows * CFG["B"] / dt, dt, kind, workers


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    if _have_reference_package():
        # import the reference (numba, sklearn, scipy ...: seconds) once, before the worker processes are forked
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import sigkernel  # noqa: F401
    # bound the sample: one probe step with one row of X (128 pairs) per worker process sizes the timed steps to ~3 s each
    # (the stock reference runs at a few hundred pairs/s per core: static kernel + tile() + single-threaded Cython solve)
    workers = max(1, min(cores, 32))
    _, t_probe, _, _ = cpu_gram_throughput(workers, workers)
    rows = min(CFG["A"], workers * max(1, min(8, int(3.0 / max(t_probe, 1e-3)))))
    times = []
    for i in range(args.warmup + args.steps):
        v, dt, kind, workers = cpu_gram_throughput(rows, workers)
        if i >= args.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    value = rows * CFG["B"] / t
    _, how = cpu_path_description()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{rows} rows of X x 128 columns = {rows * CFG['B']} pairs per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind,
                         "sample": f"{rows}x128 pairs per step; {how}; {workers} worker processes (one row block each, "
                                   "1 torch thread each: the reference itself has no multi-core path for the solve)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def measure_fp64_peak(skb, torch):
    """thread-level DP instructions / s of a register-resident DADD / DMUL / DFMA chain (burst)."""
    lib = skb._lib.lib
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    sink = torch.zeros(8, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rates = {}
    for name, op in (("dfma", 0), ("dadd", 1), ("dmul", 2)):
        best = 1e30
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            skb._lib.check(lib.skb_fp64_probe(op, sms * 8, 256, 40000, sink.data_ptr(), st))
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        rates[name] = sms * 8 * 256 * 40000 * 16 / (best * 1e-3)
    return rates


def _time_steps(torch, fn, steps, flush, barrier):
    """Device time of `steps` calls of fn in ms (this rank): every call sits between its own pair of CUDA events with an L2
    flush (256 MiB memset) before it; the whole sequence is enqueued first and synchronised once, so the host (Python,
    launch latency, jitter between ranks) is not in the loop."""
    evs = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs)


def bench_cfg4(skb, torch, dev, flush, steps, peak_rate):
    """BASELINE configs[3]: compute_mmd + .backward(), batch 128, len 64, dim 3, dyadic_order 1, RBF, through the public API
    (fused loss head: 3 forward solves, 2 reversed sweeps that rebuild the forward grid instead of reading a stored one)."""
    A, L, D, d = 128, 64, 3, 1
    g = torch.Generator().manual_seed(0)
    X = torch.rand((A, L, D), dtype=torch.float64, generator=g).to(dev)
    Y = torch.rand((A, L, D), dtype=torch.float64, generator=g).to(dev)
    sk = skb.SigKernel(skb.RBFKernel(0.5), d)
    grads = []

    def step():
        Xg = X.detach().requires_grad_(True)
        sk.compute_mmd(Xg, Y).backward()
        grads.append(Xg.grad)
        del grads[:-1]

    for _ in range(3):
        step()
    ms = _time_steps(torch, step, steps, flush, None) / steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    G, gp = skb.ops.sigkernel_forward_backward(X, Y, "rbf", 0.5, d, "gram")
    e1.record()
    torch.cuda.synchronize()
    # one Gram with per-point gradients (the reference's eager grad_points), timed over a few calls
    ts = []
    for _ in range(5):
        flush.zero_()
        e0.record()
        skb.ops.sigkernel_forward_backward(X, Y, "rbf", 0.5, d, "gram")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    MM = (L - 1) << d
    cells = A * A * MM * MM
    # algorithmic DP instructions: 4 per fine cell per sweep; the backward adds one sweep and ~2 per cell for
    # the product with the forward grid: forward of K_XX, K_XY (full) and K_YY (triangle), reversed sweeps of K_XX, K_XY
    dp = 4.0 * cells * 2.5 + 6.0 * cells * 2
    ctx_bytes = 2 * A * A * 2 * (MM + 1) * 8          # last row + last column of every grid, two Grams, written once
    return {"workload": "compute_mmd + backward batch=128 len=64 dim=3 dyadic_order=1 RBFKernel(0.5) fp64 (BASELINE.json configs[3])",
            "ms_per_step": ms, "gram_with_grad_points_ms": min(ts),
            "mmd_grad_finite": bool(torch.isfinite(grads[-1]).all().item()),
            "hbm_GB_moved": (2 * ctx_bytes + 3 * A * A * 8 + 4 * A * L * D * 8) / 1e9,
            "hbm_note": "algorithmic: boundary context written by the two forward solves and read by the two reversed sweeps, the "
                        "three Gram matrices, paths and gradient; the stored-grid kernels of round 1 moved 8.4 GB here",
            "roofline": {"bound": "fp64", "achieved": dp / (ms * 1e-3) / 1e12, "peak": peak_rate / 1e12, "unit": "T DP-instr/s",
                         "frac": dp / (ms * 1e-3) / peak_rate,
                         "formula": "(4 * 2.5 + 6 * 2) * A*B*MM*NN / t / peak  (SURVEY.md 8(d))"}}


def bench_cfg5(skb, torch, dist, dev, world, rank, flush, steps):
    """BASELINE configs[4]: compute_Gram 512x512, len 128, dim 8, dyadic_order 2, RBF, strong-scaled over the ranks with
    sigkernel_b200.distributed.compute_Gram_sharded; the leading block is checked against the reference fixture."""
    A, L, D, d = 512, 128, 8, 2
    g = torch.Generator().manual_seed(0)
    X = torch.rand((A, L, D), dtype=torch.float64, generator=g).to(dev)
    Y = torch.rand((A, L, D), dtype=torch.float64, generator=g).to(dev)
    sk = skb.SigKernel(skb.RBFKernel(0.5), d)
    out = []

    def step():
        out.append(skb.distributed.compute_Gram_sharded(sk, X, Y))
        del out[:-1]

    for _ in range(2):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    ms = _time_steps(torch, step, steps, flush, None)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    err = None
    fix = os.path.join(ROOT, "tests", "golden", "cfg5_gram_rbf.npz")
    if rank == 0 and os.path.exists(fix):
        import numpy as np
        z = np.load(fix, allow_pickle=False)
        n = z["G"].shape[0]
        got = out[-1][:n, :n].cpu().numpy()
        err = float(np.max(np.abs(got - z["G"]) / (np.abs(z["G"]) + 1.0)))
    return {"workload": "compute_Gram 512x512 len=128 dim=8 dyadic_order=2 RBFKernel(0.5) fp64 (BASELINE.json configs[4]), "
                        f"rows of X sharded over {world} ranks, strong scaling",
            "ms_per_step": ms, "pairs_per_s": A * A / (ms * 1e-3), "scaling": "strong",
            "leading_block_max_err_vs_reference_fixture": err, "tolerance": 1e-10}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import sigkernel_b200 as skb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    A, B, L, D, d = CFG["A"], CFG["B"], CFG["L"], CFG["D"], CFG["d"]
    # every rank holds all rows of X (they are small); rank r solves rows [r*A, (r+1)*A)
    Xs = [make_inputs(r, torch)[0] for r in range(world)]
    Yh = make_inputs(0, torch)[1].pin_memory()
    Xh = torch.cat(Xs, dim=0).pin_memory()
    Xd, Yd = Xh.to(dev), Yh.to(dev)
    sk = skb.SigKernel(skb.RBFKernel(CFG["sigma"]), d)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    Gh = torch.empty((world * A, B), dtype=torch.float64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gram(x, y):
        if world > 1:
            return skb.distributed.compute_Gram_sharded(sk, x, y)      # the package's sharded path (rows of X per rank)
        return sk.compute_Gram(x, y)                                     # the public, reference-shaped API

    def step_device():
        return gram(Xd, Yd)

    def step_e2e():
        x = Xh.to(dev, non_blocking=True)
        y = Yh.to(dev, non_blocking=True)
        G = gram(x, y)
        Gh.copy_(G, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return Gh

    peak = measure_fp64_peak(skb, torch) if rank == 0 else None

    # ---- kernel-resident throughput: `value` ------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                 # samples cover warm-up + timed region (both under load)
        time.sleep(0.15)
    t_load0 = time.perf_counter()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()                                           # all ranks enter the timed region together
    # EXACTLY `steps` steps, each between its own CUDA events (and the solver kernel between a second pair, recorded by the
    # library around its launch), L2 flushed before each; enqueued as one sequence and synchronised once at the end
    evs, kevs = [], []
    for _ in range(args.steps):
        flush.zero_()                                   # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(); k1.record()                        # materialise the handles
        skb._lib.lib.skb_set_profile_events(k0.cuda_event, k1.cuda_event)
        e0.record()
        step_device()
        e1.record()
        evs.append((e0, e1))
        kevs.append((k0, k1))
    torch.cuda.synchronize()
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kevs)
    t_wall1 = time.perf_counter()
    skb._lib.lib.skb_set_profile_events(None, None)
    barrier()
    clocks = sampler.stop(t_load0, t_wall1) if rank == 0 else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    pairs_per_step = world * A * B
    value = pairs_per_step * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API with host buffers: `e2e` -------------------------------
    for _ in range(max(args.warmup, 3)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_e2e = float(t.item())
    barrier()

    # ---- the other BASELINE configs that have a GPU side ------------------------------------------
    sub_steps = max(3, min(args.steps, 20))
    cfg5 = bench_cfg5(skb, torch, dist, dev, world, rank, flush, sub_steps) if world > 1 else None
    cfg4 = None
    if world == 1:
        peak_rate0 = max(peak["dadd"], peak["dmul"])
        cfg4 = bench_cfg4(skb, torch, dev, flush, sub_steps, peak_rate0)

    if rank == 0:
        k_ms = kernel_ms / args.steps
        MM = (L - 1) << d
        w_survey = 4.0 * A * B * MM * MM                        # 4 DP instructions per fine cell, nothing else
        w_full = dp_instr_per_pair(L, D, d) * A * B             # what the formulation of skb_fwd5.cuh issues (3 per cell + static kernel)
        peak_rate = max(peak["dadd"], peak["dmul"])
        prof = os.path.join(ROOT, "profiles", "r02_fwd5_cfg3_summary.json")
        traffic = None
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except (ValueError, OSError):
                traffic = None
        roofline = {
            "bound": "fp64", "achieved": w_survey / (k_ms * 1e-3) / 1e12, "peak": peak_rate / 1e12, "unit": "T DP-instr/s",
            "frac": w_survey / (k_ms * 1e-3) / peak_rate, "traffic": traffic,
            "formula": "4 * A*B*MM*NN / kernel time / peak  (SURVEY.md 8(d): stencil only, 4 DP instructions per fine cell)",
            "kernel": "fwd5_kernel<RBF,RC=4,LOGD=2,DP2=3,NW=1,LPP=16>", "kernel_ms": k_ms,
            "peak_source": "measured live: register-resident DADD/DMUL chain (skb_fp64_probe); "
                           "MEASURED_PEAKS.json has no fp64 entry",
            "peak_dfma": peak["dfma"] / 1e12,
            # the kernel's own accounting: 3 DP per cell (DADD, DMUL, DFMA) + coefficients + the static kernel
            "frac_issued_dp": w_full / (k_ms * 1e-3) / peak_rate,
            "frac_stencil_3_per_cell": 0.75 * w_survey / (k_ms * 1e-3) / peak_rate,
            "dp_instr_per_pair_issued": dp_instr_per_pair(L, D, d),
            "hbm": {"algorithmic_bytes": 8 * (A * L * D + B * L * D + A * B),
                    "achieved_GBps": 8 * (A * L * D + B * L * D + A * B) / (k_ms * 1e-3) / 1e9,
                    "peak_GBps": _hbm_peak()},
        }
        rows = 16
        try:
            if world > 1:
                raise RuntimeError("the CPU baseline is timed at N = 1 only")
            import torch as _t
            threads = _t.get_num_threads()
            v_cpu, dt_cpu, kind, workers = cpu_gram_throughput(rows, 1, threads)
            _, how = cpu_path_description()
            cpu = {"value": v_cpu, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": f"{rows} rows of X x 128 columns = {rows * B} pairs in {dt_cpu:.1f} s; {how}; one process, "
                             f"{threads} torch threads for the static kernel and tile(), the PDE solve single-threaded as in the reference"}
        except Exception as exc:  # the checker is test infrastructure; its absence must not kill the bench
            cpu = None if world > 1 else {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(exc)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": A * B, "pairs_per_step": pairs_per_step,
                       "api": ("sigkernel_b200.distributed.compute_Gram_sharded(SigKernel(RBFKernel(0.5), 2), X, Y): rows of X per "
                               "rank, Y replicated, G reassembled on every rank") if world > 1 else
                              "SigKernel(RBFKernel(0.5), 2).compute_Gram(X, Y) on device-resident tensors",
                       "l2": "flushed (256 MiB memset) between timed iterations; inputs are 0.5 MB",
                       "timing": "CUDA events around every step, steps enqueued back to back, one synchronise at the end, max over ranks"},
            "clocks": clocks,
            "e2e": {"value": pairs_per_step * args.steps / t_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": Xh.numel() * 8 + Yh.numel() * 8, "d2h_bytes_per_step": Gh.numel() * 8,
                    "ms_per_step": t_e2e / args.steps * 1e3,
                    "api": "the same call on pinned host buffers: X.cuda(), Y.cuda(), compute_Gram, .cpu()"},
            "gpu_launches": 2 * args.steps,          # per step and rank: prep2_kernel (paths + queue reset) and fwd5_kernel
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        if cfg4 is not None:
            line["cfg4"] = cfg4
        if cfg5 is not None:
            line["cfg5_sharded"] = cfg5
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError, ValueError):
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
