// skb_inst_recon5_lin_nw.cu -- fwd5_kernel in the modes of the adjoint by reconstruction (MODE_FWD_EMIT, MODE_REV_RECON;
// skb_fwd5.cuh), static kind LIN, shape group nw (see skb_recon5_launch.cuh)
#define SKB_RECON5_KIND KIND_LINEAR
#define SKB_RECON5_PART 2
#define SKB_RECON5_FN launch_group_recon5_lin_nw
#include "skb_recon5_launch.cuh"
