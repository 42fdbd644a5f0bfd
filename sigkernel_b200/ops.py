"""Operator-level wrappers: torch CUDA tensors in, torch CUDA tensors out, through the C ABI.

These mirror the reference's L3 -> L1 operator calls (sigkernel/sigkernel.py:224-234, 370-382,
440-467 and cython_backend.pyx:7, 64): PyTorch is used for device memory and streams only.
Every function requires CUDA tensors; there is no CPU path (the reference's Cython branch is the
parity oracle of this project, not a product path).
"""
import torch

from . import _lib
from ._lib import lib, check

_PAIRS = {"gram": _lib.PAIRS_GRAM, "batch": _lib.PAIRS_BATCH, "sym": _lib.PAIRS_SYM}
_STATIC = {"linear": _lib.STATIC_LINEAR, "rbf": _lib.STATIC_RBF}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _io(X, Y):
    if not (X.is_cuda and Y.is_cuda):
        raise _lib.SigKernelB200Error("sigkernel_b200 runs on CUDA tensors only (no CPU fallback); move X, Y to a GPU")
    if X.dtype != Y.dtype or X.dtype not in (torch.float64, torch.float32):
        raise _lib.SigKernelB200Error(f"X and Y must both be float64 or float32, got {X.dtype}, {Y.dtype}")
    if X.dim() != 3 or Y.dim() != 3 or X.shape[2] != Y.shape[2]:
        raise _lib.SigKernelB200Error(f"expected (batch, length, dim) paths with equal dim, got {tuple(X.shape)}, {tuple(Y.shape)}")
    return X.detach().contiguous(), Y.detach().contiguous(), (_lib.F64 if X.dtype == torch.float64 else _lib.F32)


def _workspace(nbytes, device):
    """Caller-owned scratch for one call (torch's caching allocator makes this cheap)."""
    if nbytes == 0:
        raise _lib.SigKernelB200Error("sigkernel_b200: shape not supported by this entry point "
                                      "(the backward needs ceil(len_x/32) * 2^dyadic_order <= 32)")
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def _n_out(A, B, pairs):
    return A if pairs == "batch" else A * B


def sigkernel_forward(X, Y, static_kind, static_param, dyadic_order, pairs="gram", naive=False, exact=False):
    """k(X_a, Y_b) for the pair set: (A,B) for 'gram'/'sym', (A,) for 'batch'.  fp64 result."""
    Xc, Yc, dt = _io(X, Y)
    A, M, D = Xc.shape
    B, N, _ = Yc.shape
    with torch.cuda.device(Xc.device):
        out = torch.empty(_n_out(A, B, pairs), dtype=torch.float64, device=Xc.device)
        ws, nbytes = _workspace(lib.skb_fwd_workspace_bytes(A, B, M, N, D, int(dyadic_order), _PAIRS[pairs]), Xc.device)
        check(lib.skb_sigkernel_fwd(Xc.data_ptr(), Yc.data_ptr(), dt, A, B, M, N, D, int(dyadic_order),
                                    _STATIC[static_kind], float(static_param),
                                    _lib.SCHEME_S1 if naive else _lib.SCHEME_S2, _PAIRS[pairs],
                                    _lib.ARITH_EXACT if exact else _lib.ARITH_FMA,
                                    out.data_ptr(), ws.data_ptr(), nbytes, _stream()))
    return out if pairs == "batch" else out.view(A, B)


def sigkernel_forward_peers(X, Y, static_kind, static_param, dyadic_order, peer_ptrs, pairs="gram", naive=False):
    """Forward whose results are stored to every address of `peer_ptrs` (device pointers laid out like the (A,B) /
    (A,) output: this rank's block inside each rank's copy of G, see distributed.compute_Gram_sharded) instead of a local
    tensor.  Returns False when the shape is outside the kernels that have this output path."""
    import ctypes
    Xc, Yc, dt = _io(X, Y)
    A, M, D = Xc.shape
    B, N, _ = Yc.shape
    if lib.skb_forward_plan(M, N, D, int(dyadic_order), _STATIC[static_kind], _lib.SCHEME_S1 if naive else _lib.SCHEME_S2) < 4:
        return False
    arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(q) for q in peer_ptrs])
    with torch.cuda.device(Xc.device):
        ws, nbytes = _workspace(lib.skb_fwd_workspace_bytes(A, B, M, N, D, int(dyadic_order), _PAIRS[pairs]), Xc.device)
        rc = lib.skb_sigkernel_fwd_peers(Xc.data_ptr(), Yc.data_ptr(), dt, A, B, M, N, D, int(dyadic_order),
                                         _STATIC[static_kind], float(static_param),
                                         _lib.SCHEME_S1 if naive else _lib.SCHEME_S2, _PAIRS[pairs],
                                         ctypes.cast(arr, ctypes.c_void_p), len(peer_ptrs), ws.data_ptr(), nbytes, _stream())
        if rc == -4:
            return False
        check(rc)
    return True


def static_gram(X, Y, static_kind, static_param, pairs="gram"):
    """kappa(X_a[i], Y_b[j]) for the two built-in static kernels: (A,B,M,N) fp64 ('gram') or (A,M,N) ('batch')."""
    Xc, Yc, dt = _io(X, Y)
    A, M, D = Xc.shape
    B, N, _ = Yc.shape
    with torch.cuda.device(Xc.device):
        Ks = torch.empty((A, M, N) if pairs == "batch" else (A, B, M, N), dtype=torch.float64, device=Xc.device)
        ws, nbytes = _workspace(lib.skb_fwd_workspace_bytes(A, B, M, N, D, 0, _PAIRS[pairs]), Xc.device)
        check(lib.skb_static_gram(Xc.data_ptr(), Yc.data_ptr(), dt, A, B, M, N, D, _STATIC[static_kind], float(static_param),
                                  _PAIRS[pairs], Ks.data_ptr(), ws.data_ptr(), nbytes, _stream()))
    return Ks


def n_jobs(A, B, pairs):
    """Length of the pair enumeration of `pairs` (include/sigkernel_b200.h, skb_sigkernel_fwd_range)."""
    return A if pairs == "batch" else (A * (A + 1) // 2 if pairs == "sym" else A * B)


def sigkernel_forward_range(X, Y, static_kind, static_param, dyadic_order, job_lo, job_hi, pairs="gram", out=None,
                            peer_ptrs=None, naive=False, signal=None):
    """Jobs [job_lo, job_hi) of the pair enumeration of `pairs` -- one rank's share of a sharded Gram matrix -- written
    into `out` ((A,B) fp64, other entries untouched) or to every rank's copy (`peer_ptrs`: device pointers to the START of
    each copy).  'sym' writes both mirror entries.  signal = (pointers to every rank's signal slots, this rank, epoch): the
    solver kernel ends with the barrier across the ranks (include/sigkernel_b200.h).  Returns False when the shape is
    outside the kernels with this path."""
    import ctypes
    Xc, Yc, dt = _io(X, Y)
    A, M, D = Xc.shape
    B, N, _ = Yc.shape
    if lib.skb_forward_plan(M, N, D, int(dyadic_order), _STATIC[static_kind], _lib.SCHEME_S1 if naive else _lib.SCHEME_S2) < 4:
        return False
    if (out is None) == (peer_ptrs is None):
        raise ValueError("exactly one of out / peer_ptrs")
    if out is not None and (out.dtype != torch.float64 or not out.is_contiguous() or out.numel() != (A if pairs == "batch" else A * B)):
        raise ValueError("out must be a contiguous fp64 tensor laid out like the full result")
    arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(q) for q in peer_ptrs]) if peer_ptrs is not None else None
    sarr, srank, sepoch = None, 0, 0
    if signal is not None:
        sarr = (ctypes.c_void_p * len(signal[0]))(*[int(q) for q in signal[0]])
        srank, sepoch = int(signal[1]), int(signal[2])
    with torch.cuda.device(Xc.device):
        ws, nbytes = _workspace(lib.skb_fwd_workspace_bytes(A, B, M, N, D, int(dyadic_order), _PAIRS[pairs]), Xc.device)
        rc = lib.skb_sigkernel_fwd_range(Xc.data_ptr(), Yc.data_ptr(), dt, A, B, M, N, D, int(dyadic_order),
                                         _STATIC[static_kind], float(static_param),
                                         _lib.SCHEME_S1 if naive else _lib.SCHEME_S2, _PAIRS[pairs], int(job_lo), int(job_hi),
                                         out.data_ptr() if out is not None else None,
                                         ctypes.cast(arr, ctypes.c_void_p) if arr is not None else None,
                                         len(peer_ptrs) if peer_ptrs is not None else 0,
                                         ctypes.cast(sarr, ctypes.c_void_p) if sarr is not None else None, srank, sepoch,
                                         ws.data_ptr(), nbytes, _stream())
        if rc == -4:
            return False
        check(rc)
    return True


def sigkernel_forward_from_static(Ks, dyadic_order, pairs="gram", naive=False, exact=False):
    """Plugin path: Ks is the coarse static matrix (A,B,M,N) ('gram'/'sym') or (A,M,N) ('batch')."""
    if not Ks.is_cuda:
        raise _lib.SigKernelB200Error("sigkernel_b200 runs on CUDA tensors only (no CPU fallback)")
    Kc = Ks.detach().to(torch.float64).contiguous()
    if pairs == "batch":
        A, M, N = Kc.shape
        B = A
    else:
        A, B, M, N = Kc.shape
    with torch.cuda.device(Kc.device):
        out = torch.empty(_n_out(A, B, pairs), dtype=torch.float64, device=Kc.device)
        ws, nbytes = _workspace(lib.skb_aux_workspace_bytes(A, B, M, N, int(dyadic_order), _PAIRS[pairs]), Kc.device)
        check(lib.skb_sigkernel_fwd_from_static(Kc.data_ptr(), A, B, M, N, int(dyadic_order),
                                                _lib.SCHEME_S1 if naive else _lib.SCHEME_S2, _PAIRS[pairs],
                                                _lib.ARITH_EXACT if exact else _lib.ARITH_FMA,
                                                out.data_ptr(), ws.data_ptr(), nbytes, _stream()))
    return out if pairs == "batch" else out.view(A, B)


def solve_increments(inc, naive=False, exact=True):
    """inc (..., MM, NN) fine increments -> u[MM,NN] per leading index (operator-level entry point)."""
    if not inc.is_cuda:
        raise _lib.SigKernelB200Error("sigkernel_b200 runs on CUDA tensors only (no CPU fallback)")
    lead = inc.shape[:-2]
    MM, NN = inc.shape[-2:]
    ic = inc.detach().to(torch.float64).contiguous().view(-1, MM, NN)
    P = ic.shape[0]
    with torch.cuda.device(ic.device):
        out = torch.empty(P, dtype=torch.float64, device=ic.device)
        ws, nbytes = _workspace(lib.skb_aux_workspace_bytes(P, P, MM + 1, NN + 1, 0, _lib.PAIRS_BATCH), ic.device)
        check(lib.skb_sigkernel_solve_increments(ic.data_ptr(), P, MM, NN,
                                                 _lib.SCHEME_S1 if naive else _lib.SCHEME_S2,
                                                 _lib.ARITH_EXACT if exact else _lib.ARITH_FMA,
                                                 out.data_ptr(), ws.data_ptr(), nbytes, _stream()))
    return out.view(lead)


def sigkernel_forward_backward(X, Y, static_kind, static_param, dyadic_order, pairs="gram", naive=False):
    """(k, grad_points): grad_points is (A,B,M,D) for 'gram'/'sym', (A,M,D) for 'batch' (fp64)."""
    Xc, Yc, dt = _io(X, Y)
    A, M, D = Xc.shape
    B, N, _ = Yc.shape
    with torch.cuda.device(Xc.device):
        n = _n_out(A, B, pairs)
        out = torch.empty(n, dtype=torch.float64, device=Xc.device)
        gp = torch.empty((n, M, D), dtype=torch.float64, device=Xc.device)
        ws, nbytes = _workspace(lib.skb_bwd_workspace_bytes(A, B, M, N, D, int(dyadic_order), _PAIRS[pairs]), Xc.device)
        check(lib.skb_sigkernel_fwd_bwd(Xc.data_ptr(), Yc.data_ptr(), dt, A, B, M, N, D, int(dyadic_order),
                                        _STATIC[static_kind], float(static_param),
                                        _lib.SCHEME_S1 if naive else _lib.SCHEME_S2, _PAIRS[pairs],
                                        out.data_ptr(), gp.data_ptr(), ws.data_ptr(), nbytes, _stream()))
    if pairs == "batch":
        return out, gp
    return out.view(A, B), gp.view(A, B, M, D)


def adjoint_plan(M, N, D, dyadic_order, static_kind, naive=False):
    """Kernel family the backward of this shape takes (include/sigkernel_b200.h: 6 = adjoint by reconstruction)."""
    return lib.skb_adjoint_plan(int(M), int(N), int(D), int(dyadic_order), _STATIC[static_kind],
                                _lib.SCHEME_S1 if naive else _lib.SCHEME_S2)


def adjoint_sym_supported(M, D, dyadic_order, static_kind, naive=False):
    """True if the backward of Gram(X, X) can run one reversed sweep per UNORDERED pair (sigkernel_backward_vjp, pairs 'sym')."""
    return bool(lib.skb_adjoint_sym_supported(int(M), int(D), int(dyadic_order), _STATIC[static_kind],
                                              _lib.SCHEME_S1 if naive else _lib.SCHEME_S2))


def sigkernel_forward_ctx(X, Y, static_kind, static_param, dyadic_order, pairs="gram", naive=False):
    """Forward values plus the boundary context the lazy backward needs (last row / column of every pair's grid).
    Returns (k, ctx) or None when the shape is outside the reconstruction kernels (use sigkernel_forward_backward)."""
    Xc, Yc, dt = _io(X, Y)
    A, M, D = Xc.shape
    B, N, _ = Yc.shape
    if adjoint_plan(M, N, D, dyadic_order, static_kind, naive) != 6:
        return None
    with torch.cuda.device(Xc.device):
        out = torch.empty(_n_out(A, B, pairs), dtype=torch.float64, device=Xc.device)
        cb = lib.skb_ctx_bytes(A, B, M, N, int(dyadic_order), _PAIRS[pairs])
        ctx = torch.empty(cb, dtype=torch.uint8, device=Xc.device)
        ws, nbytes = _workspace(lib.skb_fwd_workspace_bytes(A, B, M, N, D, int(dyadic_order), _PAIRS[pairs]), Xc.device)
        rc = lib.skb_sigkernel_fwd_ctx(Xc.data_ptr(), Yc.data_ptr(), dt, A, B, M, N, D, int(dyadic_order),
                                       _STATIC[static_kind], float(static_param),
                                       _lib.SCHEME_S1 if naive else _lib.SCHEME_S2, _PAIRS[pairs],
                                       out.data_ptr(), ctx.data_ptr(), cb, ws.data_ptr(), nbytes, _stream())
        if rc == -4:
            return None
        check(rc)
    return (out if pairs == "batch" else out.view(A, B)), ctx


def sigkernel_backward_vjp(X, Y, static_kind, static_param, dyadic_order, pairs, ctx, ctx_pairs, grad_out=None,
                           w_diag=0.0, w_off=0.0, out_scale=1.0, out_scale_dev=None, into=None, want_points=False,
                           naive=False):
    """pairs 'sym' (Y is X, ctx_pairs 'sym'): the same sum from one sweep per unordered pair (the sweep of (a,b) also yields
    the term of (b,a)).  Otherwise:
    Reversed sweep over every ordered pair of `pairs` ('gram' / 'batch') contracted with d loss / d K on the fly:
    gradX (A,M,D) fp64 = [into +] out_scale * [out_scale_dev] * sum_b coef(a,b) d k(X_a,Y_b)/d X_a, coef = grad_out[a,b]
    or (w_diag on a == b, w_off elsewhere).  `ctx` comes from sigkernel_forward_ctx(..., ctx_pairs).  The (A,B,M,D)
    tensor of the reference (sigkernel.py:405-416) is only materialised when want_points."""
    Xc, Yc, dt = _io(X, Y)
    A, M, D = Xc.shape
    B, N, _ = Yc.shape
    with torch.cuda.device(Xc.device):
        gradX = into if into is not None else torch.empty((A, M, D), dtype=torch.float64, device=Xc.device)
        gp = torch.empty((_n_out(A, B, pairs), M, D), dtype=torch.float64, device=Xc.device) if want_points else None
        go = None
        if grad_out is not None:
            go = grad_out.detach().to(torch.float64).contiguous()
        osd = None
        if out_scale_dev is not None:
            osd = out_scale_dev.detach().to(torch.float64).contiguous()
        ws, nbytes = _workspace(lib.skb_bwd_vjp_workspace_bytes(A, B, M, N, D, int(dyadic_order), _PAIRS[pairs]), Xc.device)
        check(lib.skb_sigkernel_bwd_vjp(Xc.data_ptr(), Yc.data_ptr(), dt, A, B, M, N, D, int(dyadic_order),
                                        _STATIC[static_kind], float(static_param),
                                        _lib.SCHEME_S1 if naive else _lib.SCHEME_S2, _PAIRS[pairs],
                                        ctx.data_ptr(), _PAIRS[ctx_pairs], go.data_ptr() if go is not None else None,
                                        float(w_diag), float(w_off), float(out_scale),
                                        osd.data_ptr() if osd is not None else None, 1 if into is not None else 0,
                                        gradX.data_ptr(), gp.data_ptr() if gp is not None else None,
                                        ws.data_ptr(), nbytes, _stream()))
    if want_points:
        return gradX, (gp if pairs == "batch" else gp.view(A, B, M, D))
    return gradX


def gram_weighted_sum(G, pairs, w_diag, w_off, acc=None):
    """acc (1,) fp64 [+]= sum_{a,b} w(a,b) G[a,b] (w_diag on the diagonal / on every batch entry, w_off elsewhere)."""
    Gc = G.detach()
    if Gc.dtype != torch.float64 or not Gc.is_contiguous():
        Gc = Gc.to(torch.float64).contiguous()
    if pairs == "batch":
        A, B = Gc.shape[0], Gc.shape[0]
    else:
        A, B = Gc.shape
    with torch.cuda.device(Gc.device):
        accumulate = acc is not None
        if acc is None:
            acc = torch.empty(1, dtype=torch.float64, device=Gc.device)
        check(lib.skb_gram_weighted_sum(Gc.data_ptr(), A, B, _PAIRS[pairs], float(w_diag), float(w_off), acc.data_ptr(),
                                        1 if accumulate else 0, _stream()))
    return acc


def sensitivity_from_static(Ks, dyadic_order, pairs="gram", naive=False):
    """Plugin path of the backward: (k, S) with S the coarse sensitivities (pairs, M-1, N-1)."""
    if not Ks.is_cuda:
        raise _lib.SigKernelB200Error("sigkernel_b200 runs on CUDA tensors only (no CPU fallback)")
    Kc = Ks.detach().to(torch.float64).contiguous()
    if pairs == "batch":
        A, M, N = Kc.shape
        B = A
    else:
        A, B, M, N = Kc.shape
    with torch.cuda.device(Kc.device):
        n = _n_out(A, B, pairs)
        out = torch.empty(n, dtype=torch.float64, device=Kc.device)
        S = torch.empty((n, M - 1, N - 1), dtype=torch.float64, device=Kc.device)
        ws, nbytes = _workspace(lib.skb_sensitivity_workspace_bytes(A, B, M, N, int(dyadic_order), _PAIRS[pairs]), Kc.device)
        check(lib.skb_sigkernel_sensitivity_from_static(Kc.data_ptr(), A, B, M, N, int(dyadic_order),
                                                        _lib.SCHEME_S1 if naive else _lib.SCHEME_S2,
                                                        _PAIRS[pairs], out.data_ptr(), S.data_ptr(),
                                                        ws.data_ptr(), nbytes, _stream()))
    if pairs == "batch":
        return out, S
    return out.view(A, B), S.view(A, B, M - 1, N - 1)


def kernel_and_derivatives_from_static(K0, K1, K2, dyadic_order, eps):
    """K0, K1, K2 (A,B,M,N): Gram_matrix(X, Y), Gram_matrix(X + eps*gamma, Y), Gram_matrix(X + 2*eps*gamma, Y)
    -> (k, k_gamma, k_gamma_gamma), each (A,B) fp64 (reference k_kgrad, sigkernel.py:504-593)."""
    if not (K0.is_cuda and K1.is_cuda and K2.is_cuda):
        raise _lib.SigKernelB200Error("sigkernel_b200 runs on CUDA tensors only (no CPU fallback)")
    if not (K0.shape == K1.shape == K2.shape) or K0.dim() != 4:
        raise _lib.SigKernelB200Error(f"expected three (A,B,M,N) static matrices, got {tuple(K0.shape)}, {tuple(K1.shape)}, {tuple(K2.shape)}")
    Ks = [k.detach().to(torch.float64).contiguous() for k in (K0, K1, K2)]
    A, B, M, N = Ks[0].shape
    with torch.cuda.device(Ks[0].device):
        out = torch.empty((A * B, 3), dtype=torch.float64, device=Ks[0].device)
        ws, nbytes = _workspace(lib.skb_deriv_workspace_bytes(A, B, M, N), Ks[0].device)
        check(lib.skb_sigkernel_derivatives_from_static(Ks[0].data_ptr(), Ks[1].data_ptr(), Ks[2].data_ptr(), A, B, M, N,
                                                        int(dyadic_order), float(eps), out.data_ptr(), ws.data_ptr(),
                                                        nbytes, _stream()))
    out = out.view(A, B, 3)
    return out[..., 0].contiguous(), out[..., 1].contiguous(), out[..., 2].contiguous()
