// skb_generic_adj.cu -- the backward of shapes outside every register-resident adjoint kernel (any len_x, len_y and
// dyadic order): the reference's own algebra on materialised grids (sigkernel.py:419-502 -- forward grid, grid of the
// reversed paths, their product pooled over the dyadic cells, static-kernel derivative), with the analytic d k / d x
// instead of the reference's finite difference.  One block per (pair, direction) walks the anti-diagonals of its grid
// in global memory; nothing here is tuned -- it exists so that no shape is refused.
#include "skb_common.cuh"
#include "skb_host.h"

namespace skb {

// U[job][dir][(MMf+1) x (NNf+1)]: dir 0 = the PDE of (x, y), dir 1 = of the reversed paths (increments flipped in both
// axes, sigkernel.py:434-438).  inc[job][Mc][Nc] = coarse increments times 4^-d.
__global__ void __launch_bounds__(256) grid_solve_kernel(const double* __restrict__ inc, double* __restrict__ U, double* __restrict__ out,
                                                          long job0, int Mc, int Nc, int d, int s1) {
    const long jl = blockIdx.x >> 1;
    const int dir = blockIdx.x & 1;
    const long MMf = (long)Mc << d, NNf = (long)Nc << d;
    const double* g = inc + jl * ((long)Mc * Nc);
    double* u = U + (2 * jl + dir) * ((MMf + 1) * (NNf + 1));
    const long W = NNf + 1;
    for (long dg = 0; dg <= MMf + NNf; ++dg) {
        const long ilo = dg > NNf ? dg - NNf : 0, ihi = dg < MMf ? dg : MMf;
        for (long i = ilo + threadIdx.x; i <= ihi; i += blockDim.x) {
            const long j = dg - i;
            double v = 1.0;
            if (i > 0 && j > 0) {
                long ci = (i - 1) >> d, cj = (j - 1) >> d;
                if (dir) { ci = Mc - 1 - ci; cj = Nc - 1 - cj; }
                const double e = g[ci * Nc + cj];
                const double u10 = u[(i - 1) * W + j], u01 = u[i * W + j - 1], u00 = u[(i - 1) * W + j - 1];
                if (s1) {
                    v = (u10 + u01) * (1.0 + 0.5 * e) - u00;          // _naive_solver (cython_backend.pyx:27)
                } else {
                    const double e12 = e * e * (1.0 / 12.0);
                    v = (u10 + u01) * (1.0 + 0.5 * e + e12) - u00 * (1.0 - e12);
                }
            }
            u[i * W + j] = v;
        }
        __syncthreads();
    }
    if (dir == 0 && threadIdx.x == 0 && out) out[job0 + jl] = u[MMf * W + NNf];
}

// S[job][i][j] = 4^-d * sum over the fine cells (p, q) of coarse cell (i, j) of u[p, q] * u_rev[MMf-1-p, NNf-1-q]
__global__ void coarse_sens_kernel(const double* __restrict__ U, double* __restrict__ S, long njobs, int Mc, int Nc, int d, double scale4) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per = (long)Mc * Nc;
    if (idx >= njobs * per) return;
    const long jl = idx / per;
    const int ij = (int)(idx - jl * per);
    const int i = ij / Nc, j = ij - i * Nc;
    const long MMf = (long)Mc << d, NNf = (long)Nc << d, W = NNf + 1;
    const double* uf = U + (2 * jl) * ((MMf + 1) * W);
    const double* ur = uf + (MMf + 1) * W;
    const int n = 1 << d;
    double s = 0.0;
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            const long p = ((long)i << d) + a, q = ((long)j << d) + b;
            s = fma(uf[p * W + q], ur[(MMf - 1 - p) * W + (NNf - 1 - q)], s);
        }
    S[idx] = s * scale4;
}

// grad[pair][p][k] = sum_j' T[p][j'] d k(x_p, y_j') / d x_pk,  T[p][j'] = dS[p][j'-1] - dS[p][j'],  dS[p][j] = S[p-1][j] - S[p][j]
// (zero outside the grid).  RBF: d k / d x = (2/sigma) (y - x) k  ->  gscale * sum W y_k - (cx x_k) sum W, W = T k;  Linear: gscale sum T y_k.
__global__ void grad_from_sens_kernel(const double* __restrict__ S, const double* __restrict__ Ks, const double* __restrict__ Xp,
                                      const double* __restrict__ Yp, double* __restrict__ grad, long job0, long njobs, int B, int M, int N,
                                      int D, int Dp, int rbf, int batch, double gscale) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per = (long)M * D;
    if (idx >= njobs * per) return;
    const long jl = idx / per;
    const int pk = (int)(idx - jl * per);
    const int p = pk / D, k = pk - p * D;
    const long pi = job0 + jl;
    const long a = batch ? pi : pi / B, b = batch ? pi : pi - a * B;
    const int Mc = M - 1, Nc = N - 1;
    const double* s = S + jl * ((long)Mc * Nc);
    const double* ks = Ks + jl * ((long)M * N) + (long)p * N;
    const double* y = Yp + b * (long)N * Dp;
    double sW = 0.0, gy = 0.0, dprev = 0.0;
    for (int j = 0; j < N; ++j) {
        double dcur = 0.0;                                   // dS[p][j], zero for j = N - 1
        if (j < Nc) dcur = (p >= 1 ? s[(long)(p - 1) * Nc + j] : 0.0) - (p < Mc ? s[(long)p * Nc + j] : 0.0);
        const double T = dprev - dcur;
        const double Wt = rbf ? T * ks[j] : T;
        sW += Wt;
        gy = fma(Wt, y[(long)j * Dp + 1 + k], gy);
        dprev = dcur;
    }
    const double xk = Xp[(a * M + p) * (long)Dp + 1 + k];
    grad[pi * per + pk] = rbf ? fma(gscale, gy, -(xk * sW)) : gscale * gy;
}

int launch_grid_solve(const double* inc, double* U, double* out, long job0, long njobs, int M, int N, int d, bool s1, cudaStream_t st) {
    if (njobs == 0) return SKB_OK;
    grid_solve_kernel<<<(unsigned)(2 * njobs), 256, 0, st>>>(inc, U, out, job0, M - 1, N - 1, d, s1 ? 1 : 0);
    return check_launch();
}

int launch_coarse_sens(const double* U, double* S, long njobs, int M, int N, int d, double scale4, cudaStream_t st) {
    const long n = njobs * (long)(M - 1) * (N - 1);
    if (n == 0) return SKB_OK;
    coarse_sens_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(U, S, njobs, M - 1, N - 1, d, scale4);
    return check_launch();
}

int launch_grad_from_sens(const double* S, const double* Ks, const KArgs& a, int kind, double* grad, long job0, long njobs, cudaStream_t st) {
    const long n = njobs * (long)a.M * a.D;
    if (n == 0) return SKB_OK;
    grad_from_sens_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S, Ks, a.Xp, a.Yp, grad, job0, njobs, a.B, a.M, a.N, a.D, a.Dp,
                                                                     kind == KIND_RBF, a.pairs == PAIRS_BATCH, a.gscale);
    return check_launch();
}

}  // namespace skb
