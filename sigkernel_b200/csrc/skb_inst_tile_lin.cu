// skb_inst_tile_lin.cu -- instantiations of tile_fwd_kernel (skb_tile.cuh), static kind Linear
#include "skb_tile_launch.cuh"

namespace skb {
int launch_group_tile_lin(int rc, int logd, int dp2, const TArgs& a, cudaStream_t st) {
    return launch_tile_group<KIND_LINEAR>(rc, logd, dp2, a, st);
}
}  // namespace skb
