"""Shared helpers for the parity tests (oracle = checker only)."""
import glob
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# Forward tolerance stated by BASELINE.json north_star / SURVEY.md 8(c):
#   |G - G_ref| <= 1e-10 * (|G_ref| + 1)
FWD_TOL = 1e-10
# Gradient tolerance vs the reference (limited by the reference's own h = 1e-9 one-sided
# finite difference) and vs the analytic restatement (oracle #2).
GRAD_TOL_REF = 5e-6
GRAD_TOL_ANALYTIC = 1e-9


def golden_names(ops=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        z = np.load(p, allow_pickle=False)
        meta = json.loads(str(z["meta"]))
        if ops is None or meta["op"] in ops:
            out.append(meta["name"])
    return out


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    arrs = {k: z[k] for k in z.files if k != "meta"}
    return meta, arrs


def fwd_err(got, ref):
    """max |got-ref| / (|ref|+1): the metric of SURVEY.md 8(c)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if got.size == 0:
        return 0.0
    return float(np.max(np.abs(got - ref) / (np.abs(ref) + 1.0)))


def grad_err(got, ref):
    """max-norm relative error, the metric the gradient tolerances were measured in."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300))


def make_paths(kind, seed, shape, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    if kind == "rand":
        return torch.rand(shape, dtype=dtype, generator=g)
    if kind == "randn":
        return torch.randn(shape, dtype=dtype, generator=g)
    if kind == "bm":
        return torch.cumsum(torch.randn(shape, dtype=dtype, generator=g) / np.sqrt(shape[1]), 1)
    raise ValueError(kind)


def static_of(mod, meta):
    """Static kernel object of module `mod` (the oracle or sigkernel_b200) for a golden fixture's metadata."""
    kind = meta["static"]
    if kind == "rbf":
        return mod.RBFKernel(meta["param"])
    if kind == "linear":
        return mod.LinearKernel(meta["param"])
    if kind == "rbf_id":
        return mod.RBF_ID_Kernel(meta["params"][0])
    if kind == "linear_id":
        return mod.Linear_ID_Kernel()
    if kind == "rbf_cexp":
        return mod.RBF_CEXP_Kernel(*meta["params"])
    raise ValueError(kind)
