// skb_inst_fwd5_lin.cu -- instantiations + launcher of fwd5_kernel (skb_fwd5.cuh), static kind Linear
#include "skb_fwd5.cuh"

namespace skb {

template <int KIND, int RC, int LOGD, int DP2>
static int launch_fwd5(const KArgs& a, cudaStream_t st) {
    constexpr int MINB = 16, UNR = 3;
    int wpsm = get_warps_per_sm() > 0 ? get_warps_per_sm() : MINB;
    if (wpsm > MINB) wpsm = MINB;
    long nw = (long)sm_count() * wpsm;
    if (nw > a.njobs) nw = a.njobs;
    fwd5_kernel<KIND, RC, LOGD, DP2, MINB, UNR><<<(unsigned)nw, 32, 0, st>>>(a);
    return check_launch();
}

int launch_group_fwd5_lin(int rc, int logd, int dp2, const KArgs& a, cudaStream_t st) {
#define SKB_CASE(RC_, LD_)                                                              \
    if (rc == RC_ && logd == LD_) {                                                     \
        switch (dp2) {                                                                  \
            case 2: return launch_fwd5<KIND_LINEAR, RC_, LD_, 2>(a, st);                   \
            case 3: return launch_fwd5<KIND_LINEAR, RC_, LD_, 3>(a, st);                   \
            case 5: return launch_fwd5<KIND_LINEAR, RC_, LD_, 5>(a, st);                   \
            default: return SKB_ERR_UNSUPPORTED;                                        \
        }                                                                               \
    }
    SKB_FWD5_SHAPES(SKB_CASE)
#undef SKB_CASE
    return SKB_ERR_UNSUPPORTED;
}

}  // namespace skb
