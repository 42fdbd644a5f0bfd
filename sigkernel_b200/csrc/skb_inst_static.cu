// skb_inst_static.cu -- instantiations for caller-provided static matrices / increments
// (plugin path and operator-level entry point), FMA and EXACT arithmetic.
#include "skb_launch.cuh"

namespace skb {
int launch_group_static(int mode, int kind, int rc, int logd, int dp2, bool exact, const KArgs& a, cudaStream_t st) {
    (void)dp2;
    if (kind == KIND_INCV) {
        if (mode != MODE_FWD || logd != 0) return SKB_ERR_UNSUPPORTED;
#define SKB_CASE(RC_)                                                                             \
    if (rc == RC_) return exact ? launch_one<MODE_FWD, KIND_INCV, RC_, 0, 0, true>(a, st)         \
                                : launch_one<MODE_FWD, KIND_INCV, RC_, 0, 0, false>(a, st);
        SKB_CASE(1) SKB_CASE(2) SKB_CASE(4) SKB_CASE(8)
#undef SKB_CASE
        return SKB_ERR_UNSUPPORTED;
    }
    if (kind == KIND_INC) {
        if (mode != MODE_FWD || logd != 0) return SKB_ERR_UNSUPPORTED;
#define SKB_CASE(RC_)                                                                             \
    if (rc == RC_) return exact ? launch_one<MODE_FWD, KIND_INC, RC_, 0, 0, true>(a, st)          \
                                : launch_one<MODE_FWD, KIND_INC, RC_, 0, 0, false>(a, st);
        SKB_CASE(1) SKB_CASE(2) SKB_CASE(4) SKB_CASE(8)
#undef SKB_CASE
        return SKB_ERR_UNSUPPORTED;
    }
#define SKB_CASE(RC_, LD_)                                                                        \
    if (rc == RC_ && logd == LD_) {                                                               \
        if (mode == MODE_FWD)                                                                     \
            return exact ? launch_one<MODE_FWD, KIND_STATIC, RC_, LD_, 0, true>(a, st)            \
                         : launch_one<MODE_FWD, KIND_STATIC, RC_, LD_, 0, false>(a, st);          \
        if (mode == MODE_FWD_STORE) return launch_one<MODE_FWD_STORE, KIND_STATIC, RC_, LD_, 0, false>(a, st); \
        if (mode == MODE_REV_S) return launch_one<MODE_REV_S, KIND_STATIC, RC_, LD_, 0, false>(a, st);         \
        return SKB_ERR_UNSUPPORTED;                                                               \
    }
    SKB_FOR_SHAPES(SKB_CASE)
#undef SKB_CASE
    return SKB_ERR_UNSUPPORTED;
}
}  // namespace skb
