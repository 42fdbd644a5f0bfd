// baseline/ref_cuda_standin.cu -- LABELLED STAND-IN for the reference's Numba-CUDA Gram kernel (BASELINE.md plan item 2).
//
// NOT product code and NOT the oracle: a plain CUDA-C restatement of the STRUCTURE of
// sigkernel/cuda_backend.py:121-160 (sigkernel_Gram_cuda) -- one block per (a, b) pair, one thread per grid row,
// a loop over the 2*threads-1 anti-diagonals with a block barrier per diagonal, increments read from and the
// solution written to global memory -- so that "the reference's own GPU path" has a measured denominator on the
// B200 even though numba 0.65 cannot be shown to target sm_100 here and /root/reference does not travel to the
// GPU box.  The launch shape (grid (A,B), max(MM,NN)+1 threads, solution (A,B,MM+2,NN+2) pre-filled with the
// boundary ones, increments zero-padded by one row/column for the reference's out-of-bounds read) follows
// sigkernel.py:370-382.  Timed by baseline/time_ref_standin.py together with the torch ops the reference runs in
// front of it (Gram_matrix, second difference, tile).
#include <cuda_runtime.h>

extern "C" __global__ void ref_gram_kernel(const double* __restrict__ inc, int len_x, int len_y, int n_anti_diagonals,
                                           double* sol, int naive) {
    // inc: (A, B, len_x, len_y) [zero-padded]; sol: (A, B, len_x + 1, len_y + 1)
    const long pair = (long)blockIdx.x * gridDim.y + blockIdx.y;
    const double* M_inc = inc + pair * (long)len_x * len_y;
    double* M_sol = sol + pair * (long)(len_x + 1) * (len_y + 1);
    const int I = threadIdx.x;
    for (int p = 0; p < n_anti_diagonals; ++p) {
        int J = p - I;
        J = J < 0 ? 0 : (J > len_y - 1 ? len_y - 1 : J);
        const int i = I + 1, j = J + 1;
        if (I + J == p && I < len_x && J < len_y) {
            const double g = M_inc[(long)(i - 1) * len_y + (j - 1)];
            const double k01 = M_sol[(long)(i - 1) * (len_y + 1) + j];
            const double k10 = M_sol[(long)i * (len_y + 1) + (j - 1)];
            const double k00 = M_sol[(long)(i - 1) * (len_y + 1) + (j - 1)];
            if (naive) M_sol[(long)i * (len_y + 1) + j] = (k01 + k10) * (1. + 0.5 * g) - k00;
            else M_sol[(long)i * (len_y + 1) + j] = (k01 + k10) * (1. + 0.5 * g + (1. / 12) * g * g) - k00 * (1. - (1. / 12) * g * g);
        }
        __syncthreads();
    }
}

extern "C" int ref_gram_launch(const double* inc, int A, int B, int len_x, int len_y, double* sol, int naive, void* stream) {
    const int tpb = len_x > len_y ? len_x : len_y;
    if (tpb > 1024) return -1;                                   // the reference asserts the same limit (sigkernel.py:368)
    ref_gram_kernel<<<dim3(A, B), tpb, 0, (cudaStream_t)stream>>>(inc, len_x, len_y, 2 * tpb - 1, sol, naive);
    return (int)cudaGetLastError();
}
