// skb_launch.cuh -- instantiation + launch helpers shared by the skb_inst_*.cu translation units.
#pragma once
#include "skb_solver.cuh"

#ifndef SKB_MINB_R8
#define SKB_MINB_R8 20
#endif

namespace skb {

template <int MODE, int R>
constexpr int minb_for() {
    // resident warps (= 1-warp blocks) per SM the register budget is sized for
    return (MODE == MODE_REV_S || MODE == MODE_REV_GRAD)
               ? (R <= 4 ? 16 : (R <= 8 ? 12 : (R <= 16 ? 8 : 6)))
               : (R <= 8 ? SKB_MINB_R8 : (R <= 16 ? 12 : 8));
}

template <int MODE, int KIND, int RC, int LOGD, int DP2, bool EXACT>
int launch_one(const KArgs& a, cudaStream_t st) {
    constexpr int R = RC << LOGD;
    constexpr int MINB = minb_for<MODE, R>();
    // default: at most 16 resident warps per SM (measured: more adds nothing once the register file
    // read ports are saturated, and fewer, longer-lived warps amortise the wavefront ramp better)
    int wpsm = get_warps_per_sm() > 0 ? get_warps_per_sm() : (MINB < 16 ? MINB : 16);
    if (wpsm > MINB) wpsm = MINB;
    long nw = (long)sm_count() * wpsm;
    if (nw > a.njobs) nw = a.njobs;
    size_t smem = KIND == KIND_RBF ? EXP_TAB * sizeof(double) : 0;
    auto kern = solver_kernel<MODE, KIND, RC, LOGD, DP2, EXACT, MINB>;
    if (MODE == MODE_REV_GRAD) {
        smem += (size_t)RC * (a.D + 1) * 32 * sizeof(double);
        if (smem > 200 * 1024) return SKB_ERR_UNSUPPORTED;
        if (smem > 48 * 1024) {
            int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (rc) return rc;
        }
    }
    kern<<<(unsigned)nw, 32, smem, st>>>(a);
    return check_launch();
}

// every (RC, LOGD) with R = RC << LOGD <= 32
#define SKB_FOR_SHAPES(X) \
    X(1, 0) X(1, 1) X(1, 2) X(1, 3) X(1, 4) X(1, 5) \
    X(2, 0) X(2, 1) X(2, 2) X(2, 3) X(2, 4) \
    X(4, 0) X(4, 1) X(4, 2) X(4, 3) \
    X(8, 0) X(8, 1) X(8, 2)

// fused kinds: dispatch over shape and the Dp/2 specialisations {0 (generic), 2, 3, 5}
template <int MODE, int KIND>
int launch_fused(int rc, int logd, int dp2, const KArgs& a, cudaStream_t st) {
#define SKB_CASE(RC_, LD_)                                                                     \
    if (rc == RC_ && logd == LD_) {                                                            \
        switch (dp2) {                                                                         \
            case 2: return launch_one<MODE, KIND, RC_, LD_, 2, false>(a, st);                  \
            case 3: return launch_one<MODE, KIND, RC_, LD_, 3, false>(a, st);                  \
            case 5: return launch_one<MODE, KIND, RC_, LD_, 5, false>(a, st);                  \
            default: return launch_one<MODE, KIND, RC_, LD_, 0, false>(a, st);                 \
        }                                                                                      \
    }
    SKB_FOR_SHAPES(SKB_CASE)
#undef SKB_CASE
    return SKB_ERR_UNSUPPORTED;
}

}  // namespace skb
