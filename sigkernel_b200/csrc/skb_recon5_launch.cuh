// skb_recon5_launch.cuh -- launch helpers of the MODE_FWD_EMIT / MODE_REV_RECON instantiations of fwd5_kernel.
// Included by the skb_inst_recon5_*.cu translation units, each of which defines SKB_RECON5_KIND (KIND_RBF / KIND_LINEAR),
// SKB_RECON5_PART (0: 16 lanes per pair; 1: 32 lanes per pair, one warp; 2: two and four warps per pair) and
// SKB_RECON5_FN (the name of the group launcher it exports).
#pragma once
#include "skb_fwd5.cuh"

namespace skb {

// (RC, LOGD) strips of at most 8 fine rows.  Keep in sync with recon5_plan() (skb_dispatch.cu).
#define SKB_RECON5_L32_SHAPES(X) X(1, 0) X(2, 0) X(4, 0) X(8, 0) X(1, 1) X(2, 1) X(4, 1) X(1, 2) X(2, 2) X(1, 3)
#define SKB_RECON5_L16_SHAPES(X) X(4, 0) X(8, 0) X(2, 1) X(4, 1) X(1, 2) X(2, 2) X(1, 3)
#define SKB_RECON5_NW_SHAPES(X) X(8, 0) X(4, 1) X(2, 2)

template <int KIND, int RC, int LOGD, int DP2, int MODE, int LPP, int NW>
int launch_recon5_one(const KArgs& a, cudaStream_t st) {
    constexpr int R = RC << LOGD;
    // resident blocks per SM the register budget is sized for (the reversed sweep carries two solutions)
    constexpr int MINB = MODE == MODE_REV_RECON_SYM ? (R <= 4 ? 12 : 8)
                         : MODE == MODE_REV_RECON ? (NW > 1 ? (NW == 2 ? 4 : 2) : (R <= 4 ? 12 : 8))
                                                  : (NW > 1 ? (NW == 2 ? 8 : 4) : (R <= 4 ? 16 : 12));
    // the reversed sweep with 16 lanes per pair is too much code to unroll 3x (110 KB: it stalled on instruction fetch)
    constexpr int UNR = (MODE == MODE_REV_RECON && LPP == 16) ? 1 : 3;
    int wpsm = get_warps_per_sm() > 0 ? get_warps_per_sm() / NW : MINB;
    if (wpsm > MINB) wpsm = MINB;
    if (wpsm < 1) wpsm = 1;
    long nb = (long)sm_count() * wpsm;
    const long need = LPP == 16 ? ((long)a.njobs + 1) / 2 : (long)a.njobs;      // two pair streams per warp at 16 lanes per pair
    if (nb > need) nb = need;
    size_t smem = 0;
    if (MODE == MODE_REV_RECON || MODE == MODE_REV_RECON_SYM) {
        constexpr bool GREG = RC * DP2 <= 6;
        smem = ((GREG ? 0 : (size_t)RC * (a.D + 1)) + (size_t)(R + 2)) * 32 * NW * sizeof(double);
        if (GREG) smem += (size_t)(a.fbuf_mask + 1) * (RC * 2 * DP2 + R) * 32 * NW * sizeof(double);   // parked sums + first columns
        if (MODE == MODE_REV_RECON_SYM) smem += (size_t)a.N * 2 * DP2 * sizeof(double);                // column sums of one pair
    }
    auto kern = fwd5_kernel<KIND, RC, LOGD, DP2, NW, MINB, UNR, MODE, LPP>;
    if (smem > 8 * 1024) {
        // static + dynamic shared memory may pass the 48 KB a kernel gets without opting in (wide paths on 2 / 4 warps)
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (rc) return rc;
    }
    kern<<<(unsigned)nb, 32 * NW, smem, st>>>(a);
    return check_launch();
}

template <int KIND, int RC, int LOGD, int DP2, int LPP, int NW>
int launch_recon5_sym(const KArgs& a, cudaStream_t st) {
    if constexpr (LPP == 32 && NW == 1 && RC * DP2 <= 6) return launch_recon5_one<KIND, RC, LOGD, DP2, MODE_REV_RECON_SYM, 32, 1>(a, st);
    else return SKB_ERR_UNSUPPORTED;
}

template <int KIND, int RC, int LOGD, int LPP, int NW>
int launch_recon5_shape(int mode, int dp2, const KArgs& a, cudaStream_t st) {
    if (mode == MODE_REV_RECON_SYM) {
        switch (dp2) {
            case 2: return launch_recon5_sym<KIND, RC, LOGD, 2, LPP, NW>(a, st);
            case 3: return launch_recon5_sym<KIND, RC, LOGD, 3, LPP, NW>(a, st);
            case 5: return launch_recon5_sym<KIND, RC, LOGD, 5, LPP, NW>(a, st);
            default: return SKB_ERR_UNSUPPORTED;
        }
    }
    if (mode == MODE_FWD_EMIT) {
        switch (dp2) {
            case 2: return launch_recon5_one<KIND, RC, LOGD, 2, MODE_FWD_EMIT, LPP, NW>(a, st);
            case 3: return launch_recon5_one<KIND, RC, LOGD, 3, MODE_FWD_EMIT, LPP, NW>(a, st);
            case 5: return launch_recon5_one<KIND, RC, LOGD, 5, MODE_FWD_EMIT, LPP, NW>(a, st);
            default: return SKB_ERR_UNSUPPORTED;
        }
    }
    switch (dp2) {
        case 2: return launch_recon5_one<KIND, RC, LOGD, 2, MODE_REV_RECON, LPP, NW>(a, st);
        case 3: return launch_recon5_one<KIND, RC, LOGD, 3, MODE_REV_RECON, LPP, NW>(a, st);
        case 5: return launch_recon5_one<KIND, RC, LOGD, 5, MODE_REV_RECON, LPP, NW>(a, st);
        default: return SKB_ERR_UNSUPPORTED;
    }
}

// (the shape lists are split over three translation units per static kind to parallelise the build)
int SKB_RECON5_FN(int mode, int rc, int logd, int dp2, int nw, const KArgs& a, cudaStream_t st) {
#if SKB_RECON5_PART == 0
#define SKB_CASE(RC_, LD_) if (rc == RC_ && logd == LD_) return launch_recon5_shape<SKB_RECON5_KIND, RC_, LD_, 16, 1>(mode, dp2, a, st);
    (void)nw;
    SKB_RECON5_L16_SHAPES(SKB_CASE)
#undef SKB_CASE
#elif SKB_RECON5_PART == 1
#define SKB_CASE(RC_, LD_) if (rc == RC_ && logd == LD_) return launch_recon5_shape<SKB_RECON5_KIND, RC_, LD_, 32, 1>(mode, dp2, a, st);
    (void)nw;
    SKB_RECON5_L32_SHAPES(SKB_CASE)
#undef SKB_CASE
#else
#define SKB_CASE(RC_, LD_)                                                                                                  \
    if (rc == RC_ && logd == LD_ && nw == 2) return launch_recon5_shape<SKB_RECON5_KIND, RC_, LD_, 32, 2>(mode, dp2, a, st); \
    if (rc == RC_ && logd == LD_ && nw == 4) return launch_recon5_shape<SKB_RECON5_KIND, RC_, LD_, 32, 4>(mode, dp2, a, st);
    SKB_RECON5_NW_SHAPES(SKB_CASE)
#undef SKB_CASE
#endif
    return SKB_ERR_UNSUPPORTED;
}

}  // namespace skb
