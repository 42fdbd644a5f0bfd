"""ctypes binding of libsigkernel_b200.so -- the C ABI declared in include/sigkernel_b200.h.

The library is the product: there is no CPU fallback and no alternative backend.  If the shared
object is missing (not built) the import of this module raises, loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SIGKERNEL_B200_LIB overrides the library path (used by tuning experiments only)
LIB_PATH = os.environ.get("SIGKERNEL_B200_LIB") or os.path.join(_HERE, "libsigkernel_b200.so")

# enums of include/sigkernel_b200.h
STATIC_LINEAR, STATIC_RBF = 0, 1
SCHEME_S2, SCHEME_S1 = 0, 1
PAIRS_GRAM, PAIRS_BATCH, PAIRS_SYM = 0, 1, 2
ARITH_FMA, ARITH_EXACT = 0, 1
F64, F32 = 0, 1

# every symbol the header declares: (name, restype, argtypes)
_vp, _i, _d, _sz, _l = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t, ctypes.c_long
SYMBOLS = {
    "skb_error_string": (ctypes.c_char_p, [_i]),
    "skb_last_cuda_error": (_i, []),
    "skb_version": (_i, []),
    "skb_set_warps_per_sm": (None, [_i]),
    "skb_set_tile_mode": (None, [_i]),
    "skb_set_adjoint_mode": (None, [_i]),
    "skb_ctx_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "skb_bwd_vjp_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "skb_sigkernel_fwd_ctx": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _d, _i, _i, _vp, _vp, _sz, _vp, _sz, _vp]),
    "skb_sigkernel_bwd_vjp": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _d, _i, _i, _vp, _i, _vp, _d, _d, _d, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "skb_gram_weighted_sum": (_i, [_vp, _i, _i, _i, _d, _d, _vp, _i, _vp]),
    "skb_set_profile_events": (None, [_vp, _vp]),
    "skb_set_deriv_mode": (None, [_i]),
    "skb_fp64_probe": (_i, [_i, _i, _i, _i, _vp, _vp]),
    "skb_forward_plan": (_i, [_i, _i, _i, _i, _i, _i]),
    "skb_adjoint_plan": (_i, [_i, _i, _i, _i, _i, _i]),
    "skb_adjoint_sym_supported": (_i, [_i, _i, _i, _i, _i]),
    "skb_fwd_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "skb_sigkernel_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _d, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "skb_sigkernel_fwd_peers": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _d, _i, _i, _vp, _i, _vp, _sz, _vp]),
    "skb_sigkernel_fwd_range": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _d, _i, _i, _l, _l, _vp, _vp, _i, _vp, _i, ctypes.c_ulonglong, _vp, _sz, _vp]),
    "skb_static_gram": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _d, _i, _vp, _vp, _sz, _vp]),
    "skb_aux_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "skb_sensitivity_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "skb_sigkernel_fwd_from_static": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "skb_sigkernel_solve_increments": (_i, [_vp, _l, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "skb_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "skb_sigkernel_fwd_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _d, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "skb_sigkernel_sensitivity_from_static": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "skb_deriv_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "skb_sigkernel_derivatives_from_static": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _d, _vp, _vp, _sz, _vp]),
}


class SigKernelB200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C sigkernel_b200/csrc` "
            "(or python -c 'import __graft_entry__ as g; g.build()').  sigkernel_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        msg = lib.skb_error_string(rc).decode()
        if rc == -5:
            msg += f" [cudaError_t {lib.skb_last_cuda_error()}]"
        raise SigKernelB200Error(f"sigkernel_b200: {msg} (code {rc})")
