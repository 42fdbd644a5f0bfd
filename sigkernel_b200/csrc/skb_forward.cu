// skb_forward.cu -- forward signature-kernel solve, hand-written for sm_100a.
//
// Replaces (reference crispitagorico/sigkernel @ 40a5831):
//   sigkernel/cuda_backend.py:6-49, 121-160     sigkernel_cuda / sigkernel_Gram_cuda (one block per pair,
//                                               one thread per grid row, global-memory anti-diagonals)
//   sigkernel/static_kernels.py:17-33, 42-73    Linear / RBF static kernels
//   sigkernel/sigkernel.py:362-364, 607-613     second difference + tile() (dyadic refinement)
// with ONE kernel in which nothing but the paths is read from HBM and one double per pair is written.
//
// Design (see DESIGN.md for the derivation and the roofline):
//   * one WARP solves one path pair at a time and STREAMS through its list of pairs;
//   * lane t owns RC coarse rows = R = RC * 2^d fine rows of the PDE grid, held in registers;
//   * time advances in "macro steps" of one COARSE column (2^d fine columns, fully unrolled);
//     lane t runs one macro step behind lane t-1 (a skewed wavefront), so that the only
//     inter-lane traffic is the 2^d bottom-row values of lane t-1 (shfl_up) and the static-kernel
//     values of the first node row of lane t+1 (shfl_down), both produced one step EARLIER --
//     all lanes execute the same instruction stream, no shared memory, no barriers;
//   * the static kernel k(x_i, y_j) for node column e is evaluated by each lane for its own rows
//     three macro steps before the stencil consumes it (registers kh1..kh3), so exp() latency
//     overlaps the dependent stencil chain of the same warp;
//   * when a lane finishes a pair it starts the next pair in the following macro step, so the
//     wavefront ramp (31 steps) is paid once per warp, not once per pair.
//
// fp64 throughout.  Per fine cell: 3 DP instructions (FMA form) or 4 (EXACT, reference rounding).
#include "skb_common.cuh"
#include "skb_host.h"

namespace skb {


__device__ __forceinline__ void job_decode(const FwdArgs& p, long j, int& a, int& b) {
    if (p.pairs == PAIRS_GRAM) {
        a = (int)(j / p.B);
        b = (int)(j - (long)a * p.B);
    } else if (p.pairs == PAIRS_BATCH) {
        a = b = (int)j;
    } else {  // upper triangle, row-major: row a holds (a,a) .. (a,A-1)
        int r = 0;
        long off = 0;
        while (off + (p.A - r) <= j) { off += p.A - r; ++r; }
        a = r;
        b = r + (int)(j - off);
    }
}

__device__ __forceinline__ void job_advance(const FwdArgs& p, int& a, int& b) {
    if (p.pairs == PAIRS_GRAM) {
        if (++b == p.B) { b = 0; ++a; }
    } else if (p.pairs == PAIRS_BATCH) {
        ++a; b = a;
    } else {
        if (++b == p.A) { ++a; b = a; }
    }
}

template <int RC, int LOGD, bool EXACT, int MINB>
__global__ void __launch_bounds__(32, MINB) fwd_kernel(const FwdArgs p) {
    constexpr int F = 1 << LOGD;   // fine columns per macro step
    constexpr int R = RC * F;      // fine rows per lane
    const int lane = threadIdx.x;
    const long wg = blockIdx.x;
    const long nw = gridDim.x;
    const long q = p.njobs / nw, rem = p.njobs % nw;
    const long jbegin = wg * q + (wg < rem ? wg : rem);
    const int J = (int)(q + (wg < rem ? 1 : 0));
    if (J == 0) return;

    const int N = p.N, M = p.M;
    int a, b;
    job_decode(p, jbegin, a, b);
    int pa = a, pb = b;      // pair whose stencil columns are still draining (the previous job)
    int jl = 0;              // local index of the job the production stream is in
    int e = -lane;           // production column inside the job; negative = lane not started

    double u[R];
#pragma unroll
    for (int r = 0; r < R; ++r) u[r] = 1.0;
    double bots[F];
#pragma unroll
    for (int f = 0; f < F; ++f) bots[f] = 1.0;
    double topprev = 1.0;
    double kh1[RC], kh2[RC], kh3[RC];
#pragma unroll
    for (int rc = 0; rc < RC; ++rc) kh1[rc] = kh2[rc] = kh3[rc] = 0.0;

    // per-lane row offsets (clamped to the last valid row: values of clamped rows are never used
    // by a valid cell, they only have to be readable)
    long xoff[RC];      // KIND_LINEAR/RBF: offset of row in Xp (in doubles, relative to pair row 0)
#pragma unroll
    for (int rc = 0; rc < RC; ++rc) {
        int row = lane * RC + rc;
        if (p.kind == KIND_STATIC || p.kind == KIND_INC) {
            row = row < p.Mv ? row : p.Mv - 1;
            xoff[rc] = (long)row * p.Nv;
        } else {
            row = row < M ? row : M - 1;
            xoff[rc] = (long)row * p.Dp;
        }
    }
    const long pairsz = (long)p.Mv * p.Nv;
    const double* xbase;   // Xp rows of X_a, or Ks block of the pair
    const double* ybase;   // Yp rows of Y_b
    auto set_bases = [&](int aa, int bb) {
        if (p.kind == KIND_STATIC || p.kind == KIND_INC) {
            const long pi = (p.pairs == PAIRS_BATCH) ? (long)aa : (long)aa * p.B + bb;
            xbase = p.Ks + pi * pairsz;
            ybase = nullptr;
        } else {
            xbase = p.Xp + (long)aa * M * p.Dp;
            ybase = p.Yp + (long)bb * N * p.Dp;
        }
    };
    set_bases(a, b);

    const bool s1 = p.s1 != 0;
    const long nsteps = (long)J * N + 2 + 31;

#pragma unroll 1
    for (long S = 0; S < nsteps; ++S) {
        // ---- 1. exchange values produced in the PREVIOUS macro step --------------------------
        double tops[F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const double t = shfl_up1(bots[f]);
            tops[f] = lane == 0 ? 1.0 : t;      // grid row 0 is the boundary u = 1
        }
        const double bk_c = shfl_down1(kh2[0]);   // lane+1 first row, node column c   (its kh2)
        const double bk_c1 = shfl_down1(kh1[0]);  // lane+1 first row, node column c+1 (its kh1)

        // ---- 2. produce the static kernel at node column `col` for this lane's rows ----------
        const int col = e < 0 ? 0 : e;
        double knew[RC];
        if (p.kind == KIND_RBF || p.kind == KIND_LINEAR) {
            const double* yp = ybase + (long)col * p.Dp;
            double2 yv = ldg2(yp);
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) {
                const double2 xv = ldg2(xbase + xoff[rc]);
                knew[rc] = fma(xv.y, yv.y, xv.x + yv.x);
            }
            for (int i = 2; i < p.Dp; i += 2) {
                yv = ldg2(yp + i);
#pragma unroll
                for (int rc = 0; rc < RC; ++rc) {
                    const double2 xv = ldg2(xbase + xoff[rc] + i);
                    knew[rc] = fma(xv.y, yv.y, fma(xv.x, yv.x, knew[rc]));
                }
            }
            if (p.kind == KIND_RBF) {
#pragma unroll
                for (int rc = 0; rc < RC; ++rc) knew[rc] = exp(knew[rc]);
            }
        } else {
            const int cc = col < p.Nv ? col : p.Nv - 1;
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) knew[rc] = __ldg(xbase + xoff[rc] + cc);
        }

        // ---- 3. stencil coefficients of coarse column c = e-3 (node columns c: kh3, c+1: kh2) -
        double ca[RC], cb[RC];
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) {
            double g;
            if (p.kind == KIND_INC) {
                g = kh3[rc];
            } else {
                const double k00 = kh3[rc], k01 = kh2[rc];
                const double k10 = rc + 1 < RC ? kh3[rc + 1 < RC ? rc + 1 : rc] : bk_c;
                const double k11 = rc + 1 < RC ? kh2[rc + 1 < RC ? rc + 1 : rc] : bk_c1;
                // ((K[i+1,j+1] + K[i,j]) - K[i+1,j]) - K[i,j+1]   (sigkernel.py:363), then / 4^d
                if (EXACT) g = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(k11, k00), -k10), -k01), p.scale4);
                else g = (((k11 + k00) - k10) - k01) * p.scale4;
            }
            coeffs<EXACT>(g, s1, ca[rc], cb[rc]);
        }

        // ---- 4. the stencil: R rows x F fine columns, all in registers ------------------------
#pragma unroll
        for (int f = 0; f < F; ++f) {
            double up = tops[f];
            double diag = f == 0 ? topprev : tops[f == 0 ? 0 : f - 1];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double left = u[r];
                const double v = cell<EXACT>(left, up, diag, ca[r >> LOGD], cb[r >> LOGD]);
                diag = left;
                up = v;
                u[r] = v;
            }
            bots[f] = up;
        }
        topprev = tops[F - 1];

        // ---- 5. a pair's last coarse column (N-2) is consumed at e == 1: emit u[MM,NN] --------
        if (e == 1 && jl >= 1 && jl <= J && lane == p.tstar) {
            double res = 0.0;
#pragma unroll
            for (int rc = 0; rc < RC; ++rc)
                if (rc == p.rcstar) res = u[(rc + 1) * F - 1];
            if (p.pairs == PAIRS_BATCH) {
                p.out[pa] = res;
            } else {
                p.out[(long)pa * p.B + pb] = res;
                if (p.pairs == PAIRS_SYM) p.out[(long)pb * p.B + pa] = res;
            }
        }
        // the step that pairs the last node column of one pair with the first of the next is a
        // dummy: use it to re-arm the left boundary u[., 0] = 1
        if (e == (N == 2 ? 0 : 2)) {
#pragma unroll
            for (int r = 0; r < R; ++r) u[r] = 1.0;
            topprev = 1.0;
        }

        // ---- 6. rotate the static-kernel history, advance the production stream ---------------
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) { kh3[rc] = kh2[rc]; kh2[rc] = kh1[rc]; kh1[rc] = knew[rc]; }
        if (++e == N) {
            e = 0;
            ++jl;
            pa = a; pb = b;
            if (jl < J) { job_advance(p, a, b); set_bases(a, b); }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// prep: X (rows, D) of type T -> Xp (rows, Dp) = (nscale*|x|^2, c*x_0 .. c*x_{D-1}, 0 ...)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void prep_kernel(const T* __restrict__ X, double* __restrict__ Xp, long rows, int D, int Dp,
                            double c, double nscale) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const T* x = X + r * D;
    double* o = Xp + r * Dp;
    double n = 0.0;
    for (int k = 0; k < D; ++k) {
        const double v = (double)x[k];
        n = fma(v, v, n);
        o[1 + k] = v * c;
    }
    o[0] = n * nscale;
    for (int k = D + 1; k < Dp; ++k) o[k] = 0.0;
}

}  // namespace skb

// ------------------------------------------------------------------------------------------------
// host side: dispatch
// ------------------------------------------------------------------------------------------------
namespace skb {

static int g_warps_per_sm = 0;
void set_warps_per_sm(int w) { g_warps_per_sm = w; }

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int RC, int LOGD, bool EXACT>
static int launch_fwd_t(const FwdArgs& args, cudaStream_t st) {
    constexpr int R = RC << LOGD;
    constexpr int MINB = R <= 8 ? 16 : (R <= 16 ? 12 : 8);
    int wpsm = g_warps_per_sm > 0 ? g_warps_per_sm : MINB;
    if (wpsm > MINB) wpsm = MINB;
    long nw = (long)sm_count() * wpsm;
    if (nw > args.njobs) nw = args.njobs;
    fwd_kernel<RC, LOGD, EXACT, MINB><<<(unsigned)nw, 32, 0, st>>>(args);
    return check_launch();
}

template <bool EXACT>
static int launch_fwd_e(int rc, int logd, const FwdArgs& a, cudaStream_t st) {
#define SKB_CASE(RC_, LD_) \
    if (rc == RC_ && logd == LD_) return launch_fwd_t<RC_, LD_, EXACT>(a, st);
    SKB_CASE(1, 0) SKB_CASE(1, 1) SKB_CASE(1, 2) SKB_CASE(1, 3) SKB_CASE(1, 4) SKB_CASE(1, 5)
    SKB_CASE(2, 0) SKB_CASE(2, 1) SKB_CASE(2, 2) SKB_CASE(2, 3) SKB_CASE(2, 4)
    SKB_CASE(4, 0) SKB_CASE(4, 1) SKB_CASE(4, 2) SKB_CASE(4, 3)
    SKB_CASE(8, 0) SKB_CASE(8, 1) SKB_CASE(8, 2)
#undef SKB_CASE
    return SKB_ERR_UNSUPPORTED;
}

// rows: production rows per pair (nodes).  Picks RC in {1,2,4,8} with 32*RC >= rows.
int launch_forward(FwdArgs args, int logd, bool exact, cudaStream_t st) {
    int rc = (args.M + 31) / 32;
    int rcp = 1;
    while (rcp < rc) rcp <<= 1;
    if (rcp > 8) return SKB_ERR_UNSUPPORTED;
    while (rcp > 1 && (rcp << logd) > 32) return SKB_ERR_UNSUPPORTED;
    if ((rcp << logd) > 32) return SKB_ERR_UNSUPPORTED;
    args.tstar = (args.M - 2) / rcp;
    args.rcstar = (args.M - 2) % rcp;
    return exact ? launch_fwd_e<true>(rcp, logd, args, st) : launch_fwd_e<false>(rcp, logd, args, st);
}

int launch_prep(const void* X, int dtype, double* Xp, long rows, int D, int Dp, double c, double nscale,
                cudaStream_t st) {
    if (rows == 0) return SKB_OK;
    const int tb = 128;
    const unsigned grid = (unsigned)((rows + tb - 1) / tb);
    if (dtype == SKB_F64)
        prep_kernel<double><<<grid, tb, 0, st>>>((const double*)X, Xp, rows, D, Dp, c, nscale);
    else
        prep_kernel<float><<<grid, tb, 0, st>>>((const float*)X, Xp, rows, D, Dp, c, nscale);
    return check_launch();
}

}  // namespace skb
