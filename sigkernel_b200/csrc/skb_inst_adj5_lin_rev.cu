// skb_inst_adj5_lin_rev.cu -- adjoint-mode instantiations of fwd5_kernel (skb_fwd5.cuh): Linear, reversed sweep + gradient
#include "skb_fwd5.cuh"

namespace skb {

template <int KIND, int RC, int LOGD, int DP2>
static int launch_adj5(const KArgs& a, cudaStream_t st) {
    constexpr int MODE = MODE_REV_GRAD;
    constexpr int MINB = 12, UNR = 3;
    int wpsm = get_warps_per_sm() > 0 ? get_warps_per_sm() : MINB;
    if (wpsm > MINB) wpsm = MINB;
    long nb = (long)sm_count() * wpsm;
    if (nb > a.njobs) nb = a.njobs;
    constexpr int FR = (RC << LOGD) << LOGD;                       // F * R doubles per lane and step
    constexpr size_t stage = FR <= 32 ? (size_t)(FR <= 8 ? 4 : (FR <= 16 ? 2 : 1)) * (FR / 2) * 32 * 16 : 0;   // cp.async ring (skb_fwd5.cuh)
    const size_t smem = MODE == MODE_REV_GRAD ? (size_t)RC * (a.D + 1) * 32 * sizeof(double) + stage : 0;
    fwd5_kernel<KIND, RC, LOGD, DP2, 1, MINB, UNR, MODE><<<(unsigned)nb, 32, smem, st>>>(a);
    return check_launch();
}

int launch_group_adj5_lin_rev(int rc, int logd, int dp2, const KArgs& a, cudaStream_t st) {
#define SKB_CASE(RC_, LD_)                                                              \
    if (rc == RC_ && logd == LD_) {                                                     \
        switch (dp2) {                                                                  \
            case 2: return launch_adj5<KIND_LINEAR, RC_, LD_, 2>(a, st);                   \
            case 3: return launch_adj5<KIND_LINEAR, RC_, LD_, 3>(a, st);                   \
            case 5: return launch_adj5<KIND_LINEAR, RC_, LD_, 5>(a, st);                   \
            default: return SKB_ERR_UNSUPPORTED;                                        \
        }                                                                               \
    }
    SKB_ADJ5_SHAPES(SKB_CASE)
#undef SKB_CASE
    return SKB_ERR_UNSUPPORTED;
}

}  // namespace skb
