// skb_common.cuh -- shared device helpers of the sigkernel_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace skb {

// production kinds of the coefficient stream (see skb_forward.cu)
constexpr int KIND_LINEAR = 0;  // k = <x', y>
constexpr int KIND_RBF = 1;     // k = exp(nx + ny + <x', y>)
constexpr int KIND_STATIC = 2;  // k read from a caller-provided coarse static matrix
constexpr int KIND_INC = 3;     // increments read directly (operator-level entry point)
constexpr int KIND_INCV = 4;    // generic fallback: fine grid swept at LOGD = 0 in row bands, increments looked up
                                // in a COARSE increment matrix (row >> d, col >> d)

constexpr int PAIRS_GRAM = 0, PAIRS_BATCH = 1, PAIRS_SYM = 2;

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(FULL, v, 1); }
__device__ __forceinline__ double shfl_down1(double v) { return __shfl_down_sync(FULL, v, 1); }

__device__ __forceinline__ double2 ldg2(const double* p) {
    return __ldg(reinterpret_cast<const double2*>(p));
}

// Stencil coefficients of one coarse cell from its increment g.
//   S2 (cython_backend.pyx:30,94,116): a = 1 + g/2 + g^2/12, b = 1 - g^2/12
//   S1 (cython_backend.pyx:27,91,114): a = 1 + g/2,          b = 1
// EXACT reproduces the rounding sequence of the reference's compiled C
// ((1. + 0.5*g) + (1./12)*(g*g); 1. - (1./12)*(g*g)), no FMA contraction.
template <bool EXACT>
__device__ __forceinline__ void coeffs(double g, bool s1, double& a, double& b) {
    const double twelfth = 1.0 / 12.0;
    if (EXACT) {
        const double h = __dadd_rn(1.0, __dmul_rn(0.5, g));
        const double t = __dmul_rn(twelfth, __dmul_rn(g, g));
        a = s1 ? h : __dadd_rn(h, t);
        b = s1 ? 1.0 : __dadd_rn(1.0, -t);
    } else {
        const double h = fma(0.5, g, 1.0);
        const double t = (g * g) * twelfth;
        a = s1 ? h : h + t;
        b = s1 ? 1.0 : 1.0 - t;
    }
}

// One stencil cell: u11 from left (u10), up (u01), diag (u00).
//   EXACT: (left + up) * a - diag * b in the reference's order, 4 roundings, no FMA.
//   FMA  : fma(a, up, fma(a, left, -(b*diag))): 3 instructions, and only ONE of them on the
//          dependency chain through `up` (the value produced one row above in the same column).
template <bool EXACT>
__device__ __forceinline__ double cell(double left, double up, double diag, double a, double b) {
    if (EXACT) {
        return __dadd_rn(__dmul_rn(__dadd_rn(left, up), a), -__dmul_rn(diag, b));
    } else {
        return fma(a, up, fma(a, left, -(b * diag)));
    }
}

}  // namespace skb
