"""sigkernel_b200 -- B200-native (sm_100a) signature-kernel PDE solver, drop-in for the hot path of
crispitagorico/sigkernel: `SigKernel(static_kernel, dyadic_order).compute_kernel / compute_Gram /
compute_mmd` (+ .backward()) and the `LinearKernel` / `RBFKernel` static-kernel plugin surface.

Importing this package loads sigkernel_b200/libsigkernel_b200.so (the C ABI of
include/sigkernel_b200.h) and raises ImportError if it has not been built: there is no CPU path.
"""
from . import _lib                                   # noqa: F401  (fails loudly if the library is missing)
from ._lib import SigKernelB200Error                 # noqa: F401
from .static_kernels import (LinearKernel, RBFKernel, RBF_CEXP_Kernel, RBF_SQR_Kernel,  # noqa: F401
                             Linear_ID_Kernel, RBF_ID_Kernel, CEXP, cos_exp_kernel)
from .sigkernel import (SigKernel, _SigKernel, _SigKernelGram, hypothesis_test, SigCHSIC,  # noqa: F401
                        c_alpha, k_kgrad)
from . import ops                                    # noqa: F401
from . import distributed                            # noqa: F401

__version__ = "0.1.0"
