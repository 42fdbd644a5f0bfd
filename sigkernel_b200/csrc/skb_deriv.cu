// skb_deriv.cu -- kernel + first / second directional derivative along gamma (three coupled Goursat stencils).
//
// Replaces (reference crispitagorico/sigkernel @ 40a5831):
//   sigkernel/sigkernel.py:504-593       k_kgrad: increment build (second differences of G(X,Y), G(X+eps g,Y),
//                                        G(X+2 eps g,Y) combined as finite differences in eps) + tile() + launch
//   sigkernel/cuda_backend.py:165-223    sigkernel_derivatives_Gram_cuda: block per pair, thread per row,
//                                        three solution grids and three increment tensors in global memory
// Here: the caller passes the three COARSE static matrices (A,B,M,N) -- never the refined (A,B,MM,NN) tensors;
// a first kernel forms the three coarse increments per cell with k_kgrad's operation order, a second one sweeps
// the fine grid anti-diagonal by anti-diagonal with the nine live diagonals in shared memory (one block per
// pair) and writes the three corner values.  Per cell the arithmetic is the reference kernel's, statement by
// statement (cuda_backend.py:205-220).  This is a first, simple mapping.
#include "skb_common.cuh"
#include "skb_host.h"

namespace skb {

// inc3[pair][i][j][0..2] = (inc, inc_diff, inc_diffdiff) of coarse cell (i, j), already divided by 4^d
__global__ void deriv_increments_kernel(const double* __restrict__ K0, const double* __restrict__ K1,
                                        const double* __restrict__ K2, double* __restrict__ inc3, long pairs, int M,
                                        int N, double eps, double scale4) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Mc = M - 1, Nc = N - 1;
    const long per = (long)Mc * Nc;
    if (idx >= pairs * per) return;
    const long p = idx / per;
    const int ij = (int)(idx - p * per);
    const int i = ij / Nc, j = ij - i * Nc;
    const long o00 = p * ((long)M * N) + (long)i * N + j, o01 = o00 + 1, o10 = o00 + N, o11 = o10 + 1;
    const double ie = 1.0 / eps, ie2 = 1.0 / (eps * eps);
    // second difference in the reference's order: ((T[i+1,j+1] + T[i,j]) - T[i+1,j]) - T[i,j+1]
    auto d2 = [&](auto T) { return __dadd_rn(__dadd_rn(__dadd_rn(T(o11), T(o00)), -T(o10)), -T(o01)); };
    const double g = d2([&](long o) { return K0[o]; });
    // G_static_diff_1 = -(1/eps) G, G_static_diff_2 = (1/eps) G(X + eps gamma)            (sigkernel.py:527-531)
    const double gd = __dadd_rn(d2([&](long o) { return __dmul_rn(-ie, K0[o]); }),
                                d2([&](long o) { return __dmul_rn(ie, K1[o]); }));
    // G_static_diffdiff_1 = -(1/eps) diff_1, _2 = -(2/eps) diff_2, _3 = (1/eps^2) G(X + 2 eps gamma)   (:533-539)
    const double gdd = __dadd_rn(__dadd_rn(d2([&](long o) { return __dmul_rn(-ie, __dmul_rn(-ie, K0[o])); }),
                                           d2([&](long o) { return __dmul_rn(-(2.0 * ie), __dmul_rn(ie, K1[o])); })),
                                 d2([&](long o) { return __dmul_rn(ie2, K2[o]); }));
    inc3[3 * idx + 0] = g * scale4;
    inc3[3 * idx + 1] = gd * scale4;
    inc3[3 * idx + 2] = gdd * scale4;
}

// One block per pair.  Node (i, j) of anti-diagonal p = i + j lives at index i of the diagonal buffers;
// three rotating diagonals x three solutions in shared memory: buf[sol][slot][i], slot = p mod 3.
__global__ void deriv_sweep_kernel(const double* __restrict__ inc3, long pairs, int Mc, int Nc, int dshift,
                                   double* __restrict__ out) {
    extern __shared__ double sm[];
    const int MM = Mc << dshift, NN = Nc << dshift;
    const int ld = MM + 1;
    double* K = sm;                 // [3][ld]
    double* Kd = sm + 3 * ld;
    double* Kdd = sm + 6 * ld;
    for (long pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
        const double* g3 = inc3 + pair * ((long)Mc * Nc) * 3;
        for (int p = 0; p <= MM + NN; ++p) {
            const int cur = p % 3, p1 = (p + 2) % 3, p2 = (p + 1) % 3;   // slots of diagonals p, p-1, p-2
            const int ilo = p > NN ? p - NN : 0, ihi = p < MM ? p : MM;
            for (int i = ilo + (int)threadIdx.x; i <= ihi; i += blockDim.x) {
                const int j = p - i;
                double k, kd, kdd;
                if (i == 0 || j == 0) {
                    k = 1.0; kd = 0.0; kdd = 0.0;               // boundary (sigkernel.py:556-557; derivatives 0)
                } else {
                    const double* gg = g3 + ((long)((i - 1) >> dshift) * Nc + ((j - 1) >> dshift)) * 3;
                    const double inc = gg[0], incd = gg[1], incdd = gg[2];
                    const double k01 = K[p1 * ld + i - 1], k10 = K[p1 * ld + i], k00 = K[p2 * ld + i - 1];
                    const double k01d = Kd[p1 * ld + i - 1], k10d = Kd[p1 * ld + i], k00d = Kd[p2 * ld + i - 1];
                    const double k01dd = Kdd[p1 * ld + i - 1], k10dd = Kdd[p1 * ld + i], k00dd = Kdd[p2 * ld + i - 1];
                    // cuda_backend.py:205-220, statement by statement (no FMA contraction)
                    const double a = __dadd_rn(__dadd_rn(1.0, __dmul_rn(0.5, inc)), __dmul_rn(1.0 / 12, __dmul_rn(inc, inc)));
                    const double b = __dadd_rn(1.0, -__dmul_rn(1.0 / 12, __dmul_rn(inc, inc)));
                    k = __dadd_rn(__dmul_rn(__dadd_rn(k01, k10), a), -__dmul_rn(k00, b));
                    const double f1 = __dadd_rn(__dmul_rn(k00, incd), __dmul_rn(k00d, inc));
                    const double f2 = __dadd_rn(__dmul_rn(k01, incd), __dmul_rn(k01d, inc));
                    const double f3 = __dadd_rn(__dmul_rn(k10, incd), __dmul_rn(k10d, inc));
                    const double base_d = __dadd_rn(__dadd_rn(k01d, k10d), -k00d);
                    const double f4 = __dadd_rn(__dmul_rn(k, incd), __dmul_rn(__dadd_rn(base_d, f1), inc));
                    kd = __dadd_rn(base_d, __dmul_rn(0.25, __dadd_rn(__dadd_rn(__dadd_rn(f1, f2), f3), f4)));
                    const double g1 = __dadd_rn(__dadd_rn(__dmul_rn(k00, incdd), __dmul_rn(__dmul_rn(2.0, k00d), incd)), __dmul_rn(k00dd, inc));
                    const double g2 = __dadd_rn(__dadd_rn(__dmul_rn(k01, incdd), __dmul_rn(__dmul_rn(2.0, k01d), incd)), __dmul_rn(k01dd, inc));
                    const double g3v = __dadd_rn(__dadd_rn(__dmul_rn(k10, incdd), __dmul_rn(__dmul_rn(2.0, k10d), incd)), __dmul_rn(k10dd, inc));
                    const double base_dd = __dadd_rn(__dadd_rn(k01dd, k10dd), -k00dd);
                    const double g4 = __dadd_rn(__dadd_rn(__dmul_rn(k, incdd), __dmul_rn(__dmul_rn(2.0, kd), incd)),
                                                __dmul_rn(__dadd_rn(base_dd, g1), inc));
                    kdd = __dadd_rn(base_dd, __dmul_rn(0.25, __dadd_rn(__dadd_rn(__dadd_rn(g1, g2), g3v), g4)));
                }
                K[cur * ld + i] = k;
                Kd[cur * ld + i] = kd;
                Kdd[cur * ld + i] = kdd;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const int last = (MM + NN) % 3;
            out[pair * 3 + 0] = K[last * ld + MM];
            out[pair * 3 + 1] = Kd[last * ld + MM];
            out[pair * 3 + 2] = Kdd[last * ld + MM];
        }
        __syncthreads();
    }
}

int launch_derivatives(const double* K0, const double* K1, const double* K2, long pairs, int M, int N, int d,
                       double eps, double* inc3, double* out3, cudaStream_t st) {
    const long cells = pairs * (long)(M - 1) * (N - 1);
    if (cells == 0) return SKB_OK;
    const double scale4 = 1.0 / (double)(1ull << (2 * d));
    deriv_increments_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(K0, K1, K2, inc3, pairs, M, N, eps, scale4);
    int rc = check_launch();
    if (rc) return rc;
    const long MM = (long)(M - 1) << d;
    const size_t smem = (size_t)9 * (MM + 1) * sizeof(double);
    if (smem > 200 * 1024) return SKB_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) {
        rc = check_cuda(cudaFuncSetAttribute(deriv_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (rc) return rc;
    }
    int threads = (int)((MM + 1 + 31) / 32 * 32);
    if (threads > 512) threads = 512;
    long blocks = pairs;
    const long cap = (long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    deriv_sweep_kernel<<<(unsigned)blocks, threads, smem, st>>>(inc3, pairs, M - 1, N - 1, d, out3);
    return check_launch();
}

}  // namespace skb
