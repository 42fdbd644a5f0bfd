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

// ---- streaming variant: one warp per stream of pairs, R fine rows per lane, the lanes one COARSE column apart ----------
// Lane l owns fine rows [l R, (l+1) R) of every pair of its warp.  Per macro step it computes the R x 2^d cells of one
// coarse column of its strip from (a) the strip's previous column (registers), (b) the row above, which lane l-1 produced
// one macro step earlier (shuffles), and (c) the three increments of the coarse cells involved, fetched one macro step
// ahead.  The pairs of a warp follow each other without draining the wavefront (lane l is l coarse columns behind lane 0,
// across pair boundaries), so the skew costs 31 macro steps per WARP, not per pair.  No shared memory, no barriers.
//
// Arithmetic: the reference's three coupled updates (cuda_backend.py:205-220) with the sums over the three neighbours
// factored out.  With q = inc/4, S = k01 + k10 + k00, T = S + k, Sd = k01' + k10' + k00', V = Sd + k':
//     k   = (k01 + k10) a - k00 b
//     k'  = (1 + 2q)(k01' + k10') + (q inc - 1) k00' + (inc'/4) T + (q inc') k00
//     k'' = (1 + 2q)(k01'' + k10'') + (q inc - 1) k00'' + (inc''/4) T + (inc'/2) V + (q inc'') k00 + (inc inc'/2) k00'
// -- 19 DP instructions per cell instead of the 62 of the statement-by-statement form; the results agree with it to
// rounding (1e-13 relative), not bit for bit.  Covers (len_x - 1) 2^d <= 256 at dyadic order <= 3.
template <int R, int LOGD>
__global__ void __launch_bounds__(128, 2) deriv_stream_kernel(const double* __restrict__ inc3, long pairs, int Mc, int Nc,
                                                              double* __restrict__ out) {
    constexpr int F = 1 << LOGD;
    constexpr int NCR = R >= F ? R / F : 1;       // coarse rows of a lane's strip
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    if (warp >= pairs) return;
    const long P = (pairs - warp + nwarps - 1) / nwarps;      // pairs of this warp: warp, warp + nwarps, ...
    const int MM = Mc << LOGD;
    const int crow0 = (lane * R) >> LOGD;         // first coarse row of the strip
    const int olane = (MM - 1) / R, orow = (MM - 1) % R;   // owner of the last grid row
    const long cell3 = (long)Mc * Nc * 3;

    double k[R], kd[R], kdd[R];                   // the strip's values at the last fine column computed
    double tpv = 1.0, tpvd = 0.0, tpvdd = 0.0;    // row above the strip at that column
    double bot[F], botd[F], botdd[F];             // bottom row of the strip over the coarse column just computed
#pragma unroll
    for (int r = 0; r < R; ++r) { k[r] = 1.0; kd[r] = 0.0; kdd[r] = 0.0; }
#pragma unroll
    for (int f = 0; f < F; ++f) { bot[f] = 1.0; botd[f] = 0.0; botdd[f] = 0.0; }
    // increments of the coarse cells (this lane's coarse rows) x (the coarse column of the NEXT macro step)
    double ni[NCR], nid[NCR], nidd[NCR];
    auto fetch = [&](long pi, int jc) {
        // pi-th pair of the warp, coarse column jc; out of range (before the first / past the last pair): zeros
#pragma unroll
        for (int c = 0; c < NCR; ++c) {
            const int cr = crow0 + c;
            const bool ok = pi >= 0 && pi < P && cr < Mc;
            const double* q = inc3 + (ok ? (warp + pi * nwarps) * cell3 + ((long)cr * Nc + jc) * 3 : 0);
            ni[c] = ok ? __ldg(q) : 0.0;
            nid[c] = ok ? __ldg(q + 1) : 0.0;
            nidd[c] = ok ? __ldg(q + 2) : 0.0;
        }
    };
    // lane l starts l macro steps late: (pi, jc) = position of the NEXT macro step of this lane in its stream of pairs
    long pi = 0;                                  // (negative: before the first pair)
    int jc = 0;                                   // counts up to Nc, then the next pair begins
    if (lane > 0) { pi = -((lane - 1) / Nc + 1); jc = (int)(((long)Nc * (-pi)) - lane); }
    fetch(pi, jc);
    const long steps = P * Nc + 31;
    for (long s = 0; s < steps; ++s) {
        // row above: what lane - 1 produced one macro step ago (same pair, same coarse column); boundary for lane 0
        double top[F], topd[F], topdd[F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            top[f] = __shfl_up_sync(FULL, bot[f], 1);
            topd[f] = __shfl_up_sync(FULL, botd[f], 1);
            topdd[f] = __shfl_up_sync(FULL, botdd[f], 1);
            if (lane == 0) { top[f] = 1.0; topd[f] = 0.0; topdd[f] = 0.0; }
        }
        // this macro step's increments (fetched one step ago) -> per-coarse-cell constants; fetch the next step's
        double ca[NCR], cnb[NCR], A1[NCR], A2[NCR], qd[NCR], c1[NCR], qdd[NCR], hh[NCR], c3[NCR], c4[NCR];
#pragma unroll
        for (int c = 0; c < NCR; ++c) {
            const double inc = ni[c], incd = nid[c], incdd = nidd[c];
            const double e12 = inc * inc * (1.0 / 12), q = 0.25 * inc;
            ca[c] = (1.0 + 0.5 * inc) + e12;
            cnb[c] = e12 - 1.0;
            A1[c] = 1.0 + 0.5 * inc;
            A2[c] = fma(q, inc, -1.0);
            qd[c] = 0.25 * incd;
            c1[c] = q * incd;
            qdd[c] = 0.25 * incdd;
            hh[c] = 0.5 * incd;
            c3[c] = q * incdd;
            c4[c] = 0.5 * inc * incd;
        }
        const bool active = pi >= 0 && pi < P;
        const long pi_now = pi;
        const int jc_now = jc;
        if (++jc == Nc) { jc = 0; ++pi; }
        fetch(pi, jc);
        if (active) {
            if (jc_now == 0) {
                // a new pair: boundary column
#pragma unroll
                for (int r = 0; r < R; ++r) { k[r] = 1.0; kd[r] = 0.0; kdd[r] = 0.0; }
                tpv = 1.0; tpvd = 0.0; tpvdd = 0.0;
            }
            double up[F], upd[F], updd[F];        // the row above the one being computed, over this coarse column
#pragma unroll
            for (int f = 0; f < F; ++f) { up[f] = top[f]; upd[f] = topd[f]; updd[f] = topdd[f]; }
            double dg0 = tpv, dg0d = tpvd, dg0dd = tpvdd;     // its value one fine column to the left
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int c = R >= F ? r >> LOGD : 0;
                double lk = k[r], ld = kd[r], ldd = kdd[r];   // cell to the left
                double dk = dg0, dd = dg0d, ddd = dg0dd;       // diagonal cell
                const double nl = lk, nld = ld, nldd = ldd;    // (becomes the next row's first diagonal)
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    const double uk = up[f], ud = upd[f], udd = updd[f];
                    const double s0 = uk + lk;
                    const double kn = fma(s0, ca[c], cnb[c] * dk);
                    const double T = (s0 + dk) + kn;
                    const double sd0 = ud + ld;
                    const double kdn = fma(A1[c], sd0, fma(A2[c], dd, fma(qd[c], T, c1[c] * dk)));
                    const double V = (sd0 + dd) + kdn;
                    const double sdd0 = udd + ldd;
                    const double kddn = fma(A1[c], sdd0, fma(A2[c], ddd, fma(qdd[c], T, fma(hh[c], V, fma(c3[c], dk, c4[c] * dd)))));
                    dk = uk; dd = ud; ddd = udd;               // this cell's "up" is the next cell's diagonal
                    lk = kn; ld = kdn; ldd = kddn;
                    up[f] = kn; upd[f] = kdn; updd[f] = kddn;  // and this row is the next row's "up"
                }
                k[r] = lk; kd[r] = ld; kdd[r] = ldd;
                dg0 = nl; dg0d = nld; dg0dd = nldd;
            }
            tpv = top[F - 1]; tpvd = topd[F - 1]; tpvdd = topdd[F - 1];
#pragma unroll
            for (int f = 0; f < F; ++f) { bot[f] = up[f]; botd[f] = upd[f]; botdd[f] = updd[f]; }
            if (jc_now == Nc - 1 && lane == olane) {
                const long pair = warp + pi_now * nwarps;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (r == orow) {
                        out[pair * 3 + 0] = k[r];
                        out[pair * 3 + 1] = kd[r];
                        out[pair * 3 + 2] = kdd[r];
                    }
            }
        }
    }
}

static int g_deriv_mode = -1;                     // 0: the diagonal kernel (bit-exact) always; else streaming where it applies
void set_deriv_mode(int mode) { g_deriv_mode = mode; }

template <int R>
static int launch_deriv_stream_r(const double* inc3, long pairs, int Mc, int Nc, int d, double* out3, cudaStream_t st) {
    long warps = pairs;
    const long cap = (long)sm_count() * 8;       // two blocks of four warps per SM (the 8-row strips take ~250 registers)
    if (warps > cap) warps = cap;
    const unsigned blocks = (unsigned)((warps + 3) / 4);
    switch (d) {
        case 0: deriv_stream_kernel<R, 0><<<blocks, 128, 0, st>>>(inc3, pairs, Mc, Nc, out3); break;
        case 1: deriv_stream_kernel<R, 1><<<blocks, 128, 0, st>>>(inc3, pairs, Mc, Nc, out3); break;
        case 2: deriv_stream_kernel<R, 2><<<blocks, 128, 0, st>>>(inc3, pairs, Mc, Nc, out3); break;
        case 3: deriv_stream_kernel<R, 3><<<blocks, 128, 0, st>>>(inc3, pairs, Mc, Nc, out3); break;
        default: return SKB_ERR_UNSUPPORTED;
    }
    return check_launch();
}

bool deriv_stream_applies(int M, int d) {
    return g_deriv_mode != 0 && d <= 3 && ((long)(M - 1) << d) <= 256;
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
    if (deriv_stream_applies(M, d)) {
        const int rows = (int)((MM + 31) / 32);
        if (rows <= 1) return launch_deriv_stream_r<1>(inc3, pairs, M - 1, N - 1, d, out3, st);
        if (rows <= 2) return launch_deriv_stream_r<2>(inc3, pairs, M - 1, N - 1, d, out3, st);
        if (rows <= 4) return launch_deriv_stream_r<4>(inc3, pairs, M - 1, N - 1, d, out3, st);
        return launch_deriv_stream_r<8>(inc3, pairs, M - 1, N - 1, d, out3, st);
    }
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
