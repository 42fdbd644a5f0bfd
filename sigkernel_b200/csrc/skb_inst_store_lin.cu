// skb_inst_store_lin.cu -- instantiations of solver_kernel<MODE_FWD_STORE, KIND_LINEAR, ...> (one group per file: parallel build)
#include "skb_launch.cuh"

namespace skb {
int launch_group_store_lin(int mode, int kind, int rc, int logd, int dp2, bool exact, const KArgs& a, cudaStream_t st) {
    (void)mode; (void)kind; (void)exact;
    return launch_fused<MODE_FWD_STORE, KIND_LINEAR>(rc, logd, dp2, a, st);
}
}  // namespace skb
