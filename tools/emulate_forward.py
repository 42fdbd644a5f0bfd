"""CPU emulation of the lane/step schedule of sigkernel_b200/csrc/skb_forward.cu (fwd_kernel).

Development aid only (there is no GPU in the build container): 32 lanes are numpy vectors,
shuffles are shifts, one warp streams a list of pairs.  It mirrors the kernel statement by
statement so that index bookkeeping (lag-3 static-kernel history, dummy/reset/emit steps, clamping,
job switching) can be checked against the oracle before spending GPU time.

    python tools/emulate_forward.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sigkernel_oracle as O  # noqa: E402


def shfl_up(v):
    o = v.copy()
    o[1:] = v[:-1]
    return o


def shfl_down(v):
    o = v.copy()
    o[:-1] = v[1:]
    return o


def emulate_warp(jobs, Kfun, M, N, RC, LOGD, s1=False, inc_kind=False):
    """jobs: list of job ids; Kfun(job, rows, col) -> production values (vector over rows).
    Returns {job: u[MM,NN]}."""
    F = 1 << LOGD
    R = RC * F
    scale4 = 1.0 / 4 ** LOGD
    lane = np.arange(32)
    J = len(jobs)
    tstar, rcstar = (M - 2) // RC, (M - 2) % RC
    e = -lane.copy()
    jl = np.zeros(32, dtype=int)
    cur = np.zeros(32, dtype=int)          # index into jobs of the production stream
    prev = np.zeros(32, dtype=int)
    u = np.ones((32, R))
    bots = np.ones((32, F))
    topprev = np.ones(32)
    kh1 = np.zeros((32, RC)); kh2 = np.zeros((32, RC)); kh3 = np.zeros((32, RC))
    out = {}
    nsteps = J * N + 2 + 31
    for S in range(nsteps):
        tops = np.empty((32, F))
        for f in range(F):
            t = shfl_up(bots[:, f])
            t[0] = 1.0
            tops[:, f] = t
        bk_c = shfl_down(kh2[:, 0])
        bk_c1 = shfl_down(kh1[:, 0])
        col = np.maximum(e, 0)
        knew = np.empty((32, RC))
        for t in range(32):
            rows = np.minimum(t * RC + np.arange(RC), M - 1)
            knew[t] = Kfun(jobs[cur[t]], rows, col[t])
        ca = np.empty((32, RC)); cb = np.empty((32, RC))
        for rc in range(RC):
            if inc_kind:
                g = kh3[:, rc]
            else:
                k00, k01 = kh3[:, rc], kh2[:, rc]
                k10 = kh3[:, rc + 1] if rc + 1 < RC else bk_c
                k11 = kh2[:, rc + 1] if rc + 1 < RC else bk_c1
                g = (((k11 + k00) - k10) - k01) * scale4
            if s1:
                ca[:, rc] = 1. + 0.5 * g; cb[:, rc] = 1.0
            else:
                ca[:, rc] = (1. + 0.5 * g) + (1. / 12) * (g * g); cb[:, rc] = 1. - (1. / 12) * (g * g)
        for f in range(F):
            up = tops[:, f].copy()
            diag = topprev.copy() if f == 0 else tops[:, f - 1].copy()
            for r in range(R):
                left = u[:, r].copy()
                v = (left + up) * ca[:, r >> LOGD] - diag * cb[:, r >> LOGD]
                diag = left; up = v; u[:, r] = v
            bots[:, f] = up
        topprev = tops[:, F - 1].copy()
        for t in range(32):
            if e[t] == 1 and 1 <= jl[t] <= J and t == tstar:
                out[jobs[prev[t]]] = u[t, (rcstar + 1) * F - 1]
            if e[t] == (0 if N == 2 else 2):
                u[t, :] = 1.0; topprev[t] = 1.0
        kh3, kh2, kh1 = kh2, kh1, knew
        e = e + 1
        for t in range(32):
            if e[t] == N:
                e[t] = 0; jl[t] += 1; prev[t] = cur[t]
                if jl[t] < J:
                    cur[t] += 1
    return out


def check(M, N, d, A=3, B=2, RC=None, seed=0, s1=False):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(A, M, 3, dtype=torch.float64, generator=g)
    Y = torch.rand(B, N, 3, dtype=torch.float64, generator=g)
    sk = O.RBFKernel(0.5)
    Ks = sk.Gram_matrix(X, Y).numpy()
    ref = O.compute_Gram(X, Y, sk, d, naive=s1).numpy()
    if RC is None:
        RC = 1
        while 32 * RC < M:
            RC *= 2
    jobs = list(range(A * B))
    out = emulate_warp(jobs, lambda j, rows, col: Ks[j // B, j % B, rows, min(col, N - 1)], M, N, RC, d, s1)
    got = np.array([out[j] for j in jobs]).reshape(A, B)
    ok = np.array_equal(got, ref)
    print(f"M={M} N={N} d={d} RC={RC} s1={s1}: bitwise={ok} maxdiff={np.abs(got - ref).max():.3e}")
    return ok


def check_inc(MM, NN, P=3, seed=0):
    rng = np.random.default_rng(seed)
    inc = rng.uniform(-0.3, 0.3, size=(P, MM, NN))
    ref = O.solve_batch(inc)[:, -1, -1]
    RC = 1
    while 32 * RC < MM + 1:
        RC *= 2
    out = emulate_warp(list(range(P)),
                       lambda j, rows, col: inc[j, np.minimum(rows, MM - 1), min(col, NN - 1)],
                       MM + 1, NN + 1, RC, 0, inc_kind=True)
    got = np.array([out[j] for j in range(P)])
    ok = np.array_equal(got, ref)
    print(f"INC MM={MM} NN={NN} RC={RC}: bitwise={ok}")
    return ok


if __name__ == "__main__":
    ok = True
    ok &= check(10, 10, 0)
    ok &= check(2, 2, 2)
    ok &= check(2, 5, 1)
    ok &= check(5, 2, 1)
    ok &= check(3, 3, 0)
    ok &= check(33, 7, 1)
    ok &= check(64, 9, 2)
    ok &= check(9, 40, 1, s1=True)
    ok &= check(32, 32, 1, A=2, B=2)
    ok &= check_inc(7, 5)
    ok &= check_inc(1, 1)
    ok &= check_inc(40, 3)
    print("ALL OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
