// skb_inst_recon5_rbf_l16.cu -- fwd5_kernel in the modes of the adjoint by reconstruction (MODE_FWD_EMIT, MODE_REV_RECON;
// skb_fwd5.cuh), static kind RBF, shape group l16 (see skb_recon5_launch.cuh)
#define SKB_RECON5_KIND KIND_RBF
#define SKB_RECON5_PART 0
#define SKB_RECON5_FN launch_group_recon5_rbf_l16
#include "skb_recon5_launch.cuh"
