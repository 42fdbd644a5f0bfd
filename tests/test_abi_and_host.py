"""CPU tests (no GPU needed): the C-ABI library loads and exports every symbol the header declares,
argument validation returns the documented codes before touching CUDA, and the host-side mirror of the
reference interface (static-kernel plugin surface, API signatures, sharding arithmetic) behaves."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

from oracle import sigkernel_oracle as O
from tests._util import ROOT, make_paths

import sigkernel_b200 as skb


def _declared_functions():
    hdr = open(os.path.join(ROOT, "include", "sigkernel_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(skb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    names = _declared_functions()
    assert len(names) >= 14
    lib = ctypes.CDLL(skb._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sigkernel_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(skb._lib.SYMBOLS) == names


def test_version_and_error_strings():
    lib = skb._lib.lib
    assert lib.skb_version() >= 2
    for code in range(0, -7, -1):
        assert lib.skb_error_string(code)
    assert b"unknown" in lib.skb_error_string(-99)


def test_argument_validation_codes_without_gpu():
    lib = skb._lib.lib
    f = lib.skb_sigkernel_fwd
    assert f(None, None, 0, 2, 2, 1, 4, 2, 0, 1, 1.0, 0, 0, 0, None, None, 0, None) == -1      # M < 2
    assert f(None, None, 0, 2, 3, 4, 4, 2, 0, 1, 1.0, 0, 1, 0, None, None, 0, None) == -1      # BATCH, A != B
    assert f(None, None, 0, 2, 2, 4, 5, 2, 0, 1, 1.0, 0, 2, 0, None, None, 0, None) == -1      # SYM, M != N
    assert f(None, None, 0, 2, 2, 4, 4, 2, 0, 7, 1.0, 0, 0, 0, None, None, 0, None) == -2      # static kind
    assert f(None, None, 0, 2, 2, 4, 4, 2, 0, 1, 1.0, 3, 0, 0, None, None, 0, None) == -2      # scheme
    assert f(None, None, 0, 2, 2, 4, 4, 2, 0, 1, 1.0, 0, 0, 1, None, None, 0, None) == -2      # exact on fused
    assert f(None, None, 0, 2, 2, 4, 4, 2, 0, 1, 1.0, 0, 0, 0, None, None, 0, None) == -6      # NULL pointers
    assert f(1, 1, 0, 2, 2, 4, 4, 2, 0, 1, 1.0, 0, 0, 0, 1, None, 0, None) == -3               # workspace
    assert lib.skb_sigkernel_solve_increments(None, 0, 4, 4, 0, 0, None, None, 0, None) == -1
    assert lib.skb_sigkernel_fwd_bwd(None, None, 0, 2, 2, 4, 4, 2, 0, 1, 1.0, 0, 2, None, None, None, 0, None) == -6      # SYM is accepted
    assert lib.skb_sigkernel_fwd_bwd(None, None, 0, 2, 3, 4, 4, 2, 0, 1, 1.0, 0, 2, None, None, None, 0, None) == -1      # ... for A == B
    assert lib.skb_fp64_probe(0, 1, 512, 1, None, None) == -1


def test_workspace_sizes():
    lib = skb._lib.lib
    assert lib.skb_fwd_workspace_bytes(0, 1, 4, 4, 2, 0, 0) == 0
    w = lib.skb_fwd_workspace_bytes(128, 128, 64, 64, 5, 2, 0)
    assert w >= 2 * 128 * 64 * 6 * 8 and w % 256 == 0
    # backward by reconstruction: prepared paths + the boundary context (last row and column of every grid: 2 * 128
    # doubles per pair here) + at most 1 GiB of forward grids for the stored-grid fallback
    b = lib.skb_bwd_workspace_bytes(128, 128, 64, 64, 3, 1, 0)
    ctx = lib.skb_ctx_bytes(128, 128, 64, 64, 1, 0)
    assert ctx == 128 * 128 * 2 * 128 * 8
    assert ctx + (1 << 30) - (1 << 22) <= b <= ctx + (1 << 30) + (1 << 22)
    assert lib.skb_adjoint_plan(64, 64, 3, 1, 1, 0) == 6 and lib.skb_adjoint_plan(1000, 20, 2, 0, 0, 0) == 6
    assert lib.skb_adjoint_plan(300, 40, 2, 1, 1, 0) == 6 and lib.skb_adjoint_plan(64, 64, 3, 1, 1, 1) == 1   # S1: stored grid
    # stored-grid kernels only: one padded forward grid per pair (row pitch 32 * rows-per-lane), capped at 8 GiB
    lib.skb_set_adjoint_mode(0)
    try:
        b = lib.skb_bwd_workspace_bytes(128, 128, 64, 64, 3, 1, 0)
        assert 128 * 128 * 126 * 128 * 8 <= b <= 128 * 128 * 126 * 128 * 8 + (1 << 22)
        assert lib.skb_bwd_workspace_bytes(512, 512, 128, 128, 8, 2, 0) <= (8 << 30) + (64 << 20)
        assert lib.skb_adjoint_plan(64, 64, 3, 1, 1, 0) == 5
    finally:
        lib.skb_set_adjoint_mode(-1)
    assert lib.skb_bwd_vjp_workspace_bytes(128, 128, 64, 64, 3, 1, 0) > 0
    # beyond the register-resident kernels: materialised grids, any length (plan 7); only the eager entry point takes them
    per = (2000 * 8 + 2 * 1999 * 7 + 2 * 2000 * 8) * 8
    assert 4 * per <= lib.skb_bwd_workspace_bytes(2, 2, 2000, 8, 2, 0, 0) <= 4 * per + (1 << 20)
    assert lib.skb_adjoint_plan(2000, 8, 2, 0, 1, 0) == 7 and lib.skb_adjoint_plan(300, 9, 2, 2, 0, 0) == 7
    assert lib.skb_bwd_vjp_workspace_bytes(2, 2, 2000, 8, 2, 0, 0) == 0
    assert lib.skb_aux_workspace_bytes(2, 2, 8, 8, 1, 0) >= 4
    # shapes outside the register-resident kernels get the generic row-band workspace
    assert lib.skb_fwd_workspace_bytes(2, 2, 2000, 8, 2, 0, 0) > lib.skb_fwd_workspace_bytes(2, 2, 200, 8, 2, 0, 0)


def test_no_cpu_fallback():
    X = make_paths("rand", 0, (2, 5, 2))
    with pytest.raises(skb.SigKernelB200Error, match="CUDA"):
        skb.SigKernel(skb.RBFKernel(1.0), 0).compute_Gram(X, X)
    with pytest.raises(skb.SigKernelB200Error):
        skb.ops.solve_increments(torch.zeros(1, 3, 3, dtype=torch.float64))


def test_api_signatures_match_reference():
    """Names, positional order and defaults of the reference's public surface (sigkernel.py:18-197)."""
    sig = inspect.signature
    assert list(sig(skb.SigKernel.__init__).parameters) == ["self", "static_kernel", "dyadic_order", "_naive_solver"]
    assert sig(skb.SigKernel.__init__).parameters["_naive_solver"].default is False
    for name, params in {
        "compute_kernel": ["self", "X", "Y", "max_batch"],
        "compute_Gram": ["self", "X", "Y", "sym", "max_batch"],
        "compute_distance": ["self", "X", "Y", "max_batch"],
        "compute_scoring_rule": ["self", "X", "y", "max_batch"],
        "compute_expected_scoring_rule": ["self", "X", "Y", "max_batch"],
        "compute_mmd": ["self", "X", "Y", "max_batch"],
    }.items():
        p = sig(getattr(skb.SigKernel, name)).parameters
        assert list(p) == params
        assert p["max_batch"].default == 100
    assert sig(skb.SigKernel.compute_Gram).parameters["sym"].default is False
    assert list(sig(skb._SigKernelGram.forward).parameters) == ["ctx", "X", "Y", "static_kernel", "dyadic_order", "sym", "_naive_solver"]
    assert list(sig(skb._SigKernel.forward).parameters) == ["ctx", "X", "Y", "static_kernel", "dyadic_order", "_naive_solver"]
    assert skb.LinearKernel().scale == 1.0 and skb.RBFKernel(0.3).sigma == 0.3


def test_requires_grad_assertions_like_reference():
    sk = skb.SigKernel(skb.RBFKernel(1.0), 0)
    X = make_paths("rand", 0, (2, 5, 2))
    Y = make_paths("rand", 1, (2, 5, 2)).requires_grad_(True)
    for fn in (sk.compute_mmd, sk.compute_distance, sk.compute_scoring_rule, sk.compute_expected_scoring_rule):
        with pytest.raises(AssertionError, match="second input"):
            fn(X, Y)


def test_static_kernel_plugin_surface_matches_oracle():
    """The Python-side Gram_matrix / batch_kernel stay callable on any device (they are the plugin
    interface) and reproduce the reference arithmetic bit for bit."""
    X, Y = make_paths("randn", 2, (3, 7, 4)), make_paths("randn", 3, (3, 5, 4))
    for mine, ref in ((skb.RBFKernel(0.7), O.RBFKernel(0.7)), (skb.LinearKernel(0.5), O.LinearKernel(0.5))):
        assert torch.equal(mine.batch_kernel(X, Y), ref.batch_kernel(X, Y))
        assert torch.equal(mine.Gram_matrix(X, Y), ref.Gram_matrix(X, Y))
    assert skb.RBFKernel(0.7).fused_spec(True) == ("rbf", 0.7, None)
    assert skb.LinearKernel(0.5).fused_spec(True)[:2] == ("linear", 1.0)      # Gram ignores scale
    assert skb.LinearKernel(0.5).fused_spec(False)[:2] == ("linear", 0.25)    # batch uses scale^2


def test_function_space_kernels():
    X4, Y4 = make_paths("rand", 4, (2, 6, 3, 2)), make_paths("rand", 5, (2, 5, 3, 2))
    k = skb.RBF_ID_Kernel(0.9)
    assert torch.allclose(k.Gram_matrix(X4, Y4), O.RBFKernel(0.9).Gram_matrix(X4.reshape(2, 6, 6), Y4.reshape(2, 5, 6)))
    kind, par, tr = k.fused_spec(True)
    assert (kind, par) == ("rbf", 0.9) and tr(X4).shape == (2, 6, 6)
    lk = skb.Linear_ID_Kernel()
    assert torch.allclose(lk.batch_kernel(X4, Y4), torch.bmm(X4.reshape(2, 6, 6), Y4.reshape(2, 5, 6).transpose(1, 2)))
    ck = skb.RBF_CEXP_Kernel(1.0, 0.8, 3)
    assert ck.Gram_matrix(X4.double(), Y4.double()).shape == (2, 2, 6, 5)
    assert ck.fused_spec(True)[0] == "rbf"
    sq = skb.RBF_SQR_Kernel(1.0, 2.0)
    assert sq.batch_kernel(X4, Y4).shape == (2, 6, 5) and not hasattr(sq, "fused_spec")


def test_host_gradient_helpers_match_oracle():
    """The plugin backward's host arithmetic (finite-difference d inc/dx contracted with S) equals the
    oracle's restatement of prep_backward when fed the oracle's S."""
    from sigkernel_b200.sigkernel import _grad_points_from_sensitivity, _perturbed, _second_diff_x, _H_FD
    X, Y = make_paths("rand", 6, (2, 6, 3)), make_paths("rand", 7, (3, 5, 3))
    sk = O.RBFKernel(0.5)
    _, gp_ref = O.gram_grad_points(X, Y, sk, 1)
    U, Ks = O.gram_grid(X, Y, sk, 1)
    inc = O.increments(Ks, 1)
    Urev = torch.flip(torch.from_numpy(O.solve_gram(torch.flip(inc, dims=[2, 3]).numpy())), dims=[2, 3])
    S = O.coarse_sensitivity(U, Urev, 1)
    A, M, D = X.shape
    Kh = sk.Gram_matrix(_perturbed(X), Y).reshape(A, 3, M, D, -1).permute(0, 1, 2, 4, 3)
    hi, lo = _second_diff_x(Kh, Ks)
    gp = _grad_points_from_sensitivity(S, hi, lo) / _H_FD
    assert np.max(np.abs(gp.numpy() - gp_ref.numpy())) <= 1e-9 * np.max(np.abs(gp_ref.numpy()))


def test_dispatch_plans_for_the_baseline_configs():
    """Host-side dispatch (no GPU): every BASELINE config takes the v5 kernels; odd shapes fall back as documented."""
    lib = skb._lib.lib
    LIN, RBF, S2, S1 = 0, 1, 0, 1
    # forward: M, N, D, d, kind, scheme -> plan
    assert lib.skb_forward_plan(10, 10, 2, 0, LIN, S2) == 4      # cfg1: 16 lanes per pair
    assert lib.skb_forward_plan(32, 32, 3, 1, RBF, S2) == 4      # cfg2
    assert lib.skb_forward_plan(64, 64, 5, 2, RBF, S2) == 4      # cfg3 (headline)
    assert lib.skb_forward_plan(64, 64, 3, 1, RBF, S2) == 4      # cfg4 forward
    assert lib.skb_forward_plan(64, 64, 5, 3, RBF, S2) == 5      # dyadic order 3: one warp, 16-row strips
    assert lib.skb_forward_plan(200, 9, 3, 2, RBF, S2) == 7      # len_x = 200 at dyadic order 2: four warps per pair
    assert lib.skb_forward_plan(100, 11, 8, 1, RBF, S2) == 6     # len_x > 64: 32 lanes per pair, 2 warps
    assert lib.skb_forward_plan(128, 128, 8, 2, RBF, S2) == 5    # cfg5: one warp per pair, 16-row strips
    assert lib.skb_forward_plan(250, 9, 3, 2, RBF, S2) == 7      # four warps per pair
    assert lib.skb_forward_plan(64, 64, 5, 2, RBF, S1) == 4      # _naive_solver: an S1 instantiation of the single-warp fwd5 variants
    assert lib.skb_forward_plan(250, 9, 3, 2, RBF, S1) == 1      # ... several warps per pair: v4 kernel
    assert lib.skb_forward_plan(64, 3, 5, 2, RBF, S2) == 1       # len_y < 4: v4 kernel
    assert lib.skb_forward_plan(64, 64, 12, 2, RBF, S2) == 1     # dim + 1 > 10: generic-width v4 kernel
    assert lib.skb_forward_plan(1000, 6, 2, 0, RBF, S2) == 0     # beyond the register-resident kernels: row bands
    assert lib.skb_forward_plan(1, 6, 2, 0, RBF, S2) == -1
    assert lib.skb_forward_plan(8, 6, 2, 0, 9, S2) == -2
    # backward
    assert lib.skb_adjoint_plan(64, 64, 3, 1, RBF, S2) == 6      # cfg4: adjoint by reconstruction
    assert lib.skb_adjoint_plan(64, 64, 5, 2, RBF, S2) == 6
    assert lib.skb_adjoint_plan(40, 70, 8, 0, RBF, S2) == 6      # dyadic order 0 too
    assert lib.skb_adjoint_plan(128, 128, 8, 2, RBF, S2) == 6    # two warps per pair
    assert lib.skb_adjoint_plan(1000, 6, 2, 0, RBF, S2) == 6     # the reference's own limit: (len_x - 1) 2^d < 1024
    assert lib.skb_adjoint_plan(64, 64, 3, 1, RBF, S1) == 1      # _naive_solver: stored-grid v4 kernels
    assert lib.skb_adjoint_plan(64, 3, 3, 1, RBF, S2) == 1       # len_y < 4
    assert lib.skb_adjoint_plan(2000, 6, 2, 0, RBF, S2) == 7     # any other length: materialised grids (never unsupported)
    assert lib.skb_adjoint_plan(300, 6, 2, 2, RBF, S2) == 7      # 1196 fine rows
    lib.skb_set_adjoint_mode(0)
    try:
        assert lib.skb_adjoint_plan(64, 64, 3, 1, RBF, S2) == 5      # stored grid on the v5 kernels
        assert lib.skb_adjoint_plan(40, 70, 8, 0, RBF, S2) == 1      # dyadic order 0: v4 adjoint kernels
        assert lib.skb_adjoint_plan(128, 128, 8, 2, RBF, S2) == 1    # more than one warp per pair: v4
    finally:
        lib.skb_set_adjoint_mode(-1)
