// skb_tile_launch.cuh -- launch helpers of tile_fwd_kernel shared by the skb_inst_tile_*.cu translation units
#pragma once
#include "skb_tile.cuh"

namespace skb {

// (RC, LOGD) strips of 16 fine rows.  Keep in sync with tile_shape_ok() (skb_dispatch.cu).
#define SKB_TILE_SHAPES(X) X(4, 2) X(8, 1) X(2, 3)

constexpr int TILE_W = 8;    // warps per block = strips per band

template <int KIND, int RC, int LOGD, int DP2>
size_t tile_smem_bytes() {
    constexpr int F = 1 << LOGD, H = (F + 1) / 2, W = TILE_W, RD = tile_ring_depth(RC, DP2);
    constexpr bool XREG = (RC * DP2 <= 20);
    return (KIND == KIND_RBF ? TTAB * sizeof(double) : 0) + (size_t)(W + 1) * RD * (H + 1) * 32 * sizeof(double2) +
           (size_t)(XREG ? 1 : 2) * RC * DP2 * 32 * W * sizeof(double2) + (W + 2) * sizeof(unsigned) + TILE_JR * sizeof(int);
}

template <int KIND, int RC, int LOGD, int DP2>
int launch_tile_one(const TArgs& a, cudaStream_t st) {
    auto kern = tile_fwd_kernel<KIND, RC, LOGD, DP2, TILE_W>;
    const size_t smem = tile_smem_bytes<KIND, RC, LOGD, DP2>();
    static bool attr_done = false;       // per instantiation; idempotent, so a race between threads is harmless
    if (!attr_done) {
        int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (rc) return rc;
        attr_done = true;
    }
    long nb = sm_count();                // one block of TILE_W warps per SM: all blocks co-resident (the band hand-off
    if (nb > a.njobs) nb = a.njobs;      // between jobs relies on that)
    kern<<<(unsigned)nb, 32 * (TILE_W + 1), smem, st>>>(a);   // TILE_W stencil warps + the helper warp
    return check_launch();
}

template <int KIND>
int launch_tile_group(int rc, int logd, int dp2, const TArgs& a, cudaStream_t st) {
#define SKB_CASE(RC_, LD_)                                                              \
    if (rc == RC_ && logd == LD_) {                                                     \
        switch (dp2) {                                                                  \
            case 2: return launch_tile_one<KIND, RC_, LD_, 2>(a, st);                   \
            case 3: return launch_tile_one<KIND, RC_, LD_, 3>(a, st);                   \
            case 5: return launch_tile_one<KIND, RC_, LD_, 5>(a, st);                   \
            default: return SKB_ERR_UNSUPPORTED;                                        \
        }                                                                               \
    }
    SKB_TILE_SHAPES(SKB_CASE)
#undef SKB_CASE
    return SKB_ERR_UNSUPPORTED;
}

}  // namespace skb
