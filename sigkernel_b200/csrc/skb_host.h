// skb_host.h -- host-side declarations shared by the translation units of libsigkernel_b200.so
#pragma once
#include <cuda_runtime.h>
#include "../../include/sigkernel_b200.h"

namespace skb {

struct FwdArgs {
    const double* Xp;   // prepped X rows [A*M][Dp]: (nx, c*x_0 .. c*x_{D-1}, 0 pad)
    const double* Yp;   // prepped Y rows [B*N][Dp]: (ny,   y_0 ..   y_{D-1}, 0 pad)
    const double* Ks;   // KIND_STATIC: coarse static matrix; KIND_INC: fine increments
    double* out;
    long njobs;
    int A, B;
    int M, N;           // production rows / columns per pair (nodes; KIND_INC: MM+1, NN+1)
    int Mv, Nv;         // rows / columns actually stored in Ks
    int Dp;             // doubles per prepped row (even)
    int kind, pairs, s1;
    int tstar, rcstar;  // lane / coarse row that ends up holding u[MM,NN]
    double scale4;      // 4^-d (dyadic refinement: tile()/2^d twice, sigkernel.py:364)
};

int launch_forward(FwdArgs args, int logd, bool exact, cudaStream_t st);

// records the cudaError_t for skb_last_cuda_error(); returns SKB_OK or SKB_ERR_CUDA
int check_cuda(cudaError_t e);
inline int check_launch() { return check_cuda(cudaGetLastError()); }

void set_warps_per_sm(int w);
int launch_prep(const void* X, int dtype, double* Xp, long rows, int D, int Dp, double c, double nscale,
                cudaStream_t st);

}  // namespace skb
