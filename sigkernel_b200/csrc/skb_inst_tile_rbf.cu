// skb_inst_tile_rbf.cu -- instantiations of tile_fwd_kernel (skb_tile.cuh), static kind RBF
#include "skb_tile_launch.cuh"

namespace skb {
int launch_group_tile_rbf(int rc, int logd, int dp2, const TArgs& a, cudaStream_t st) {
    return launch_tile_group<KIND_RBF>(rc, logd, dp2, a, st);
}
}  // namespace skb
