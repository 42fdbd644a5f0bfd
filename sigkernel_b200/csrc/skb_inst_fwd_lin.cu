// skb_inst_fwd_lin.cu -- instantiations of solver_kernel<MODE_FWD, KIND_LINEAR, ...> (one group per file: parallel build)
#include "skb_launch.cuh"

namespace skb {
int launch_group_fwd_lin(int mode, int kind, int rc, int logd, int dp2, bool exact, const KArgs& a, cudaStream_t st) {
    (void)mode; (void)kind; (void)exact;
    return launch_fused<MODE_FWD, KIND_LINEAR>(rc, logd, dp2, a, st);
}
}  // namespace skb
