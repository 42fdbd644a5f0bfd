// skb_inst_fwd5_lin_l16.cu -- fwd5_kernel with 16 lanes per pair (two pair streams per warp), static kind Linear
#include "skb_fwd5.cuh"

namespace skb {

template <int KIND, int RC, int LOGD, int DP2>
static int launch_l16(const KArgs& a, cudaStream_t st) {
    constexpr int MINB = (RC << LOGD) <= 8 ? 16 : 8, UNR = 3;     // 16-row strips need ~180 registers
    int wpsm = get_warps_per_sm() > 0 ? get_warps_per_sm() : MINB;
    if (wpsm > 16) wpsm = 16;        // an explicit skb_set_warps_per_sm() may exceed the default residency (tuning)
    long nb = (long)sm_count() * wpsm;
    const long need = ((long)a.njobs + 1) / 2;                   // two pair streams per warp
    if (nb > need) nb = need;
    if (a.s1) fwd5_kernel<KIND, RC, LOGD, DP2, 1, MINB, UNR, 0, 16, true><<<(unsigned)nb, 32, 0, st>>>(a);
    else fwd5_kernel<KIND, RC, LOGD, DP2, 1, MINB, UNR, 0, 16><<<(unsigned)nb, 32, 0, st>>>(a);
    return check_launch();
}

int launch_group_fwd5_lin_l16(int rc, int logd, int dp2, const KArgs& a, cudaStream_t st) {
#define SKB_CASE(RC_, LD_)                                                              \
    if (rc == RC_ && logd == LD_) {                                                     \
        switch (dp2) {                                                                  \
            case 2: return launch_l16<KIND_LINEAR, RC_, LD_, 2>(a, st);                    \
            case 3: return launch_l16<KIND_LINEAR, RC_, LD_, 3>(a, st);                    \
            case 5: return launch_l16<KIND_LINEAR, RC_, LD_, 5>(a, st);                    \
            default: return SKB_ERR_UNSUPPORTED;                                        \
        }                                                                               \
    }
    SKB_FWD5_L16_SHAPES(SKB_CASE)
#undef SKB_CASE
    return SKB_ERR_UNSUPPORTED;
}

}  // namespace skb
