// skb_inst_fwd5_rbf.cu -- instantiations + launcher of fwd5_kernel (skb_fwd5.cuh), static kind RBF
#include "skb_fwd5.cuh"

namespace skb {

template <int KIND, int RC, int LOGD, int DP2, int NW>
static int launch_fwd5(const KArgs& a, cudaStream_t st) {
    constexpr int MINB = ((RC << LOGD) > 8 ? 8 : 16) / NW, UNR = 3;       // 16 resident warps per SM (8 with 16-row strips)
    int bpsm = get_warps_per_sm() > 0 ? get_warps_per_sm() / NW : MINB;
    if (bpsm > MINB) bpsm = MINB;
    if (bpsm < 1) bpsm = 1;
    long nb = (long)sm_count() * bpsm;
    if (nb > a.njobs) nb = a.njobs;
    if (NW == 1 && a.s1) fwd5_kernel<KIND, RC, LOGD, DP2, 1, MINB * NW, UNR, 0, 32, true><<<(unsigned)nb, 32, 0, st>>>(a);
    else fwd5_kernel<KIND, RC, LOGD, DP2, NW, MINB, UNR><<<(unsigned)nb, 32 * NW, 0, st>>>(a);
    return check_launch();
}

template <int NW>
static int launch_nw(int rc, int logd, int dp2, const KArgs& a, cudaStream_t st) {
#define SKB_CASE(RC_, LD_)                                                              \
    if (rc == RC_ && logd == LD_) {                                                     \
        switch (dp2) {                                                                  \
            case 2: return launch_fwd5<KIND_RBF, RC_, LD_, 2, NW>(a, st);               \
            case 3: return launch_fwd5<KIND_RBF, RC_, LD_, 3, NW>(a, st);               \
            case 5: return launch_fwd5<KIND_RBF, RC_, LD_, 5, NW>(a, st);               \
            default: return SKB_ERR_UNSUPPORTED;                                        \
        }                                                                               \
    }
    SKB_FWD5_SHAPES(SKB_CASE)
    SKB_FWD5_R16_SHAPES(SKB_CASE)       // this translation unit is NW = 1
#undef SKB_CASE
    return SKB_ERR_UNSUPPORTED;
}

int launch_group_fwd5_rbf_nw1(int rc, int logd, int dp2, const KArgs& a, cudaStream_t st) {
    return launch_nw<1>(rc, logd, dp2, a, st);
}

}  // namespace skb
