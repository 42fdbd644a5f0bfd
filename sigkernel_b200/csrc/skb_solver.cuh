// skb_solver.cuh -- the one kernel template behind every entry point of libsigkernel_b200.so.
//
// Replaces (reference crispitagorico/sigkernel @ 40a5831):
//   sigkernel/cuda_backend.py:6-49, 121-160     sigkernel_cuda / sigkernel_Gram_cuda (one block per pair,
//                                               one thread per grid row, global-memory anti-diagonals)
//   sigkernel/static_kernels.py:17-33, 42-73    Linear / RBF static kernels
//   sigkernel/sigkernel.py:362-364, 607-613     second difference + tile() (dyadic refinement)
//   sigkernel/sigkernel.py:419-502, 256-343     prep_backward / _SigKernel.backward (reversed PDE,
//                                               GG = u * u_rev, per-point gradients)
//
// Design (DESIGN.md has the derivation, the roofline and the measurements):
//   * one WARP solves one path pair at a time and STREAMS through pairs taken from an atomic queue;
//   * lane t owns RC coarse rows = R = RC * 2^d fine rows of the PDE grid, held in registers;
//   * time advances in "macro steps" of one COARSE column (F = 2^d fine columns, fully unrolled);
//     lane t runs one macro step behind lane t-1 (a skewed wavefront), so the only inter-lane traffic
//     is the F bottom-row values of lane t-1 (shfl_up) and the static-kernel values of the first node
//     row of lane t+1 (shfl_down), both produced at least one step EARLIER: every lane executes the
//     same instruction stream, there is no shared-memory grid, no barrier;
//   * the static kernel k(x_i, y_j) at node column e is evaluated by each lane for its own node rows
//     three macro steps before the stencil consumes it (registers kh1..kh3): exp() and the L1 latency
//     of the path loads overlap the dependent stencil chain of the same warp;
//   * a lane that finishes a pair starts the next one in the following macro step: the 31-step
//     wavefront ramp is paid once per warp, not once per pair.  Per pair there are N production steps
//     but N-1 coarse columns; the odd step out (it would pair the last node column of one pair with
//     the first of the next) re-arms the boundary u[., 0] = 1 (by select, so that an overflowed pair
//     cannot leak NaNs into the next one);
//   * fp64 instructions on this part are paced by register-operand reads: DADD/DMUL issue every 2
//     cycles per scheduler, a DFMA with three distinct register operands every 3 (measured with
//     skb_fp64_probe ops 3-6).  The cell update is therefore written as DADD + DMUL + DFMA
//     (u11 = a (u10 + u01) + (-b) u00: 7 operand-cycles) rather than three chained DFMAs (9).
//
// MODE_FWD        out[pair] = u[MM, NN]
// MODE_FWD_STORE  additionally stores u[p, q] (p < MM, q < NN) for the adjoint pass
// MODE_REV_S      same sweep on the REVERSED paths (= the reference's flipped-increment solve,
//                 sigkernel.py:438-469), multiplies by the stored forward grid and reduces to coarse
//                 sensitivities S (plugin path: written out)
// MODE_REV_GRAD   ... and contracts S with the analytic static-kernel derivative into per-point gradients
//
// fp64 throughout.  Per fine cell: 3 DP instructions (FMA form) or 4 (EXACT: reference rounding order).
#pragma once
#include "skb_common.cuh"
#include "skb_host.h"

namespace skb {

constexpr int MODE_FWD = 0, MODE_FWD_STORE = 1, MODE_REV_S = 2, MODE_REV_GRAD = 3, MODE_REV_RECON = 4, MODE_FWD_EMIT = 5,
              MODE_REV_RECON_SYM = 6;

// exp(x) for the RBF static kernel: x <= ~0 (|x - y|^2 >= 0 up to rounding), possibly hugely negative.
// Table-driven: x = (256 n + j) ln2/256 + r, exp(x) = 2^n * T[j] * e^r with T[j] = 2^(j/256) in shared
// memory and a degree-4 Taylor polynomial for |r| <= ln2/512 (truncation 4e-17 relative, <= 1 ulp
// overall).  9 DP
// instructions instead of libdevice's 16+; fp64 issue is the resource this kernel is bound by.
// Results below e^-700 flush to 0 (torch.exp returns ~1e-304 there; absolute difference < 1e-300).
constexpr int EXP_TAB = 256;
__device__ __forceinline__ void exp_table_fill(double* tab, int lane) {
    for (int j = lane; j < EXP_TAB; j += 32) tab[j] = exp2((double)j * (1.0 / EXP_TAB));
    __syncwarp();
}
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ tab) {
    const double K_L2E = 369.32993046757463;            // 256 / ln 2
    const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
    const double C_HI = 0x1.62e42fee00000p-9;           // ln2/256 = C_HI + C_LO; C_HI has 21 trailing zero bits,
    const double C_LO = 0x1.a39ef35793c76p-41;          // so n * C_HI is exact for |256 n + j| < 2^21
    const bool tiny = x < -700.0;
    const double xc = tiny ? -700.0 : x;
    const double t = fma(xc, K_L2E, MAGIC);
    const double nf = t - MAGIC;
    double r = fma(nf, -C_HI, xc);
    r = fma(nf, -C_LO, r);
    double q = fma(r, 4.1666666666666664e-2, 1.6666666666666666e-1);   // e^r - 1 = r (1 + r/2 + r^2/6 + r^3/24)
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = q * r;
    const int ti = __double2loint(t);                    // low word of t = 256 n + j (two's complement)
    const double tj = tab[ti & (EXP_TAB - 1)];
    const double v = fma(tj, q, tj);                     // <= 1 ulp overall (checked against libm)
    const double res = __hiloint2double(__double2hiint(v) + ((ti >> 8) << 20), __double2loint(v));
    return tiny ? 0.0 : res;
}

// upper triangle, row-major: row a holds (a,a) .. (a,A-1); off(r) = r (2A - r + 1) / 2 <= j < off(r + 1): closed form, then at
// most a step of correction either way.  Kept out of line: the kernels decode a job in several places of their unrolled
// loops, and the square root would be inlined into each of them (the symmetric enumeration is the rare case).
static __device__ __noinline__ int2 sym_job_decode(int A, long j) {
    const double t = 2.0 * A + 1.0;
    int r = (int)((t - sqrt(fmax(t * t - 8.0 * (double)j, 0.0))) * 0.5);
    r = r < 0 ? 0 : (r > A - 1 ? A - 1 : r);
    long off = (long)r * (2L * A - r + 1) / 2;
    while (off > j) { --r; off = (long)r * (2L * A - r + 1) / 2; }
    while (off + (A - r) <= j) { off += A - r; ++r; }
    return make_int2(r, r + (int)(j - off));
}

__device__ __forceinline__ void job_decode(const KArgs& p, long j, int& a, int& b) {
    if (p.pairs == PAIRS_GRAM) {
        a = (int)(j / p.B);
        b = (int)(j - (long)a * p.B);
    } else if (p.pairs == PAIRS_BATCH) {
        a = b = (int)j;
    } else {
        const int2 ab = sym_job_decode(p.A, j);
        a = ab.x;
        b = ab.y;
    }
}

template <int MODE, int KIND, int RC, int LOGD, int DP2, bool EXACT, int MINB>
__global__ void __launch_bounds__(32, MINB) solver_kernel(const KArgs p) {
    if (MODE != MODE_FWD && p.cond != nullptr) {
        // adjoint passes queued as the fallback of the reconstruction adjoint: run only if it raised its flag
        if (*reinterpret_cast<const volatile unsigned int*>(p.cond) == 0u) return;
    }
    constexpr int F = 1 << LOGD;   // fine columns per macro step
    constexpr int R = RC * F;      // fine rows per lane
    constexpr bool REV = (MODE == MODE_REV_S || MODE == MODE_REV_GRAD);
    constexpr bool FUSED = (KIND == KIND_RBF || KIND == KIND_LINEAR);
    constexpr bool BAND = (KIND == KIND_INCV);   // row-band sweep of the fine grid (generic fallback)
    constexpr bool VEC = (F >= 2);  // MM and R even: 16-byte aligned scratch rows
    constexpr int UNR = (MODE == MODE_FWD && FUSED && R <= 16) ? 3 : 1;   // macro steps per loop trip
    const int lane = threadIdx.x;
    const int N = p.N, M = p.M;
    const int NS = N < 3 ? 3 : N;  // macro steps per pair (N = 2 is padded with one idle production)
    const int Dp = FUSED ? (DP2 > 0 ? 2 * DP2 : p.Dp) : 0;

    extern __shared__ double smem_raw[];
    double* const etab = smem_raw;                                   // KIND_RBF: 2^(j/256) table
    double* const smem = smem_raw + (KIND == KIND_RBF ? EXP_TAB : 0);  // MODE_REV_GRAD: accumulators
                                                                     // [(rc*(D+1)+k)*32 + lane]
    if (KIND == KIND_RBF) exp_table_fill(etab, threadIdx.x);

    // ---- job stream state (per lane; lane t runs t steps behind lane 0) ----------------------------
    int job = blockIdx.x;                   // production job (local index in [0, njobs))
    int job_next = 0;                       // lane 0: prefetched next job
    if (lane == 0) job_next = (int)(gridDim.x + atomicAdd(p.counter, 1u));
    int a, b;                               // pair of the production stream
    job_decode(p, p.job0 + job, a, b);
    int sa = a, sb = b;                     // pair of the stencil stream (3 steps behind)
    int sjob = job;
    bool svalid = false;                    // stencil stream holds a real pair
    bool pvalid = true;                     // production stream holds a real pair
    int e = -lane;                          // production column; negative = lane not started
    int c = NS - 3 - lane;                  // stencil column = e - 3 (mod NS); set properly below
    while (c < 0) c += NS;                  // (value irrelevant until the first dummy step re-arms)

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

    // per-lane row offsets (clamped: values of clamped rows never reach a valid cell)
    // (fused kinds: 32-bit offsets in doubles -- the prepared paths are far below 16 GB)
    long xoff[RC];
    unsigned xoff32[RC];
#pragma unroll
    for (int rc = 0; rc < RC; ++rc) {
        int row = lane * RC + rc;
        if (BAND) {
            int frow = p.band_row0 + row;                      // global fine row
            const int fmax = (p.Mc << p.dshift) - 1;
            frow = frow < fmax ? frow : fmax;
            xoff[rc] = (long)(frow >> p.dshift) * p.Nc;
            xoff32[rc] = 0;
        } else if (!FUSED) {
            row = row < p.Mv ? row : p.Mv - 1;
            if (REV) row = p.Mv - 1 - row;
            xoff[rc] = (long)row * p.Nv;
            xoff32[rc] = 0;
        } else {
            row = row < M ? row : M - 1;
            xoff[rc] = 0;
            xoff32[rc] = (unsigned)(row * Dp);     // REV: Xp already holds the reversed path
        }
    }
    const unsigned xstride = (unsigned)(M * Dp), ystride = (unsigned)(N * Dp);
    const double* xp[RC];
    const double* yb = nullptr;
    auto set_ptrs = [&]() {
        if (FUSED) {
            const unsigned xb = (unsigned)a * xstride;
            yb = p.Yp + (unsigned)b * ystride;
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) xp[rc] = p.Xp + (xb + xoff32[rc]);
        } else {
            const long pi = (p.pairs == PAIRS_BATCH) ? (long)a : (long)a * p.B + b;
            const double* base = p.Ks + (BAND ? (long)job * ((long)p.Mc * p.Nc) : pi * ((long)p.Mv * p.Nv));
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) xp[rc] = base + xoff[rc];
        }
    };
    set_ptrs();
    const double* syb = yb;                 // Y rows of the stencil stream's pair (REV_GRAD)
    const double* sxb = FUSED ? p.Xp + (unsigned)a * xstride : nullptr;

    // REV: coarse sensitivities of this and the previous column; gradient accumulators in smem
    double Sprev[RC];
#pragma unroll
    for (int rc = 0; rc < RC; ++rc) Sprev[rc] = 0.0;
    double Slast_cur = 0.0, Slast_prev = 0.0;   // lane's LAST coarse row at the two newest columns
    const int D = p.D;
    if (MODE == MODE_REV_GRAD) {
        for (int i = 0; i < RC * (D + 1); ++i) smem[i * 32 + lane] = 0.0;
    }

    const double tw = p.s1 ? 0.0 : 1.0 / 12.0;
    const int tstar = p.tstar, rcstar = p.rcstar;
    const long NN = (long)(N - 1) << LOGD;

    // One macro step.  The loop below runs UNR of them per trip so that the compiler can
    // rename the loop-carried registers (static-kernel history, grid column) instead of moving them:
    // register-file reads, moves included, are the resource this kernel is bound by.
    auto step = [&]() __attribute__((always_inline)) {
        // ---- 1. exchange values produced in EARLIER macro steps -----------------------------------
        double tops[F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const double t = shfl_up1(bots[f]);
            tops[f] = lane == 0 ? 1.0 : t;              // grid row 0 is the boundary u = 1
            if (BAND) {
                // band > 0: the row above comes from the previous band's launch
                if (lane == 0 && p.band_top != nullptr && svalid && c < N - 1)
                    tops[f] = p.band_top[(long)sjob * (N - 1) + c];
            }
        }
        const double bk_c = shfl_down1(kh2[0]);         // lane+1 first node row, node column c   (its kh2)
        const double bk_c1 = shfl_down1(kh1[0]);        // lane+1 first node row, node column c+1 (its kh1)
        const int na = __shfl_up_sync(FULL, a, 1);      // the pair lane-1 is producing for
        const int nb_ = __shfl_up_sync(FULL, b, 1);
        const int njob = __shfl_up_sync(FULL, job, 1);
        double up_c = 0.0, up_c1 = 0.0;                 // REV: S of lane-1's last coarse row
        if (REV) {
            up_c = shfl_up1(Slast_cur);
            up_c1 = shfl_up1(Slast_prev);
            if (lane == 0) { up_c = 0.0; up_c1 = 0.0; }
        }

        // ---- 2. produce the static kernel at node column `col` for this lane's node rows ----------
        const int col = e < 0 ? 0 : (e < N ? e : N - 1);
        double knew[RC];
        if (FUSED) {
            const double* yp = yb + (long)col * Dp;
            if (DP2 > 0) {
                double2 yv[DP2 > 0 ? DP2 : 1];
#pragma unroll
                for (int i = 0; i < DP2; ++i) yv[i] = ldg2(yp + 2 * i);
#pragma unroll
                for (int rc = 0; rc < RC; ++rc) {
                    double2 xv = ldg2(xp[rc]);
                    double acc = fma(xv.y, yv[0].y, xv.x + yv[0].x);
#pragma unroll
                    for (int i = 1; i < DP2; ++i) {
                        xv = ldg2(xp[rc] + 2 * i);
                        acc = fma(xv.y, yv[i].y, fma(xv.x, yv[i].x, acc));
                    }
                    knew[rc] = acc;
                }
            } else {
                double2 yv = ldg2(yp);
#pragma unroll
                for (int rc = 0; rc < RC; ++rc) {
                    const double2 xv = ldg2(xp[rc]);
                    knew[rc] = fma(xv.y, yv.y, xv.x + yv.x);
                }
#pragma unroll 1
                for (int i = 2; i < Dp; i += 2) {
                    yv = ldg2(yp + i);
#pragma unroll
                    for (int rc = 0; rc < RC; ++rc) {
                        const double2 xv = ldg2(xp[rc] + i);
                        knew[rc] = fma(xv.y, yv.y, fma(xv.x, yv.x, knew[rc]));
                    }
                }
            }
            if (KIND == KIND_RBF) {
#pragma unroll
                for (int rc = 0; rc < RC; ++rc) knew[rc] = exp_neg(knew[rc], etab);
            }
        } else {
            int cc = col < p.Nv ? col : p.Nv - 1;
            if (REV) cc = p.Nv - 1 - cc;
            if (BAND) cc >>= p.dshift;
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) knew[rc] = __ldg(xp[rc] + cc);
        }

        // ---- 3. stencil coefficients of coarse column c (node columns c: kh3, c+1: kh2) ------------
        // FMA mode holds (a, -b); both modes re-arm u = 1 on the dummy step by select (section 4).
        const bool dummy = c >= N - 1;                   // no such coarse column: re-arm the boundary
        double ca[RC], cb[RC];
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) {
            double g;
            if (KIND == KIND_INC || KIND == KIND_INCV) {
                g = kh3[rc];
            } else {
                const double k00 = kh3[rc], k01 = kh2[rc];
                const double k10 = rc + 1 < RC ? kh3[rc + 1 < RC ? rc + 1 : rc] : bk_c;
                const double k11 = rc + 1 < RC ? kh2[rc + 1 < RC ? rc + 1 : rc] : bk_c1;
                // ((K[i+1,j+1] + K[i,j]) - K[i+1,j]) - K[i,j+1]  (sigkernel.py:363), then / 4^d.
                // In the reversed sweep (k00 <-> K[i+1,j+1] ...) the two subtractions swap so that
                // the rounding sequence of the ORIGINAL cell is reproduced.
                if (EXACT) {
                    g = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(k11, k00), -k10), -k01), p.scale4);
                } else if (REV) {
                    g = (((k11 + k00) - k01) - k10) * p.scale4;
                } else {
                    g = (((k11 + k00) - k10) - k01) * p.scale4;
                }
            }
            if (EXACT) {
                coeffs<true>(g, p.s1 != 0, ca[rc], cb[rc]);
            } else {
                // a = 1 + g/2 + g^2/12, -b = g^2/12 - 1  (tw = 0 for the S1 scheme)
                const double gg = g * g;
                ca[rc] = fma(gg, tw, fma(g, 0.5, 1.0));
                cb[rc] = fma(gg, tw, -1.0);
            }
        }

        // ---- 4. the stencil: R rows x F fine columns, all in registers -----------------------------
        const bool real_col = svalid && !dummy;
        double sacc[RC];
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) sacc[rc] = 0.0;
        double* srow = nullptr;   // scratch row of fine column q = c*F (+f), this lane's first row
        if (MODE == MODE_FWD_STORE) {
            srow = p.scratch + (((long)sjob * NN + (long)c * F) * p.pitch + (long)lane * R);
        } else if (REV) {
            // reversed coordinates: p = MM-1-(lane*R + r), q = NN-1-(c*F + f)
            const long MMl = (long)(M - 1) << LOGD;
            srow = p.scratch + (((long)sjob * NN + (NN - 1 - (long)c * F)) * p.pitch + (MMl - (long)(lane + 1) * R));
        }
        // forward values u[p, q] of the cells this lane sweeps in this step (REV: reversed row order)
        double fw[F][R];
        if (REV) {
#pragma unroll
            for (int f = 0; f < F; ++f) {
#pragma unroll
                for (int r = 0; r < R; ++r) fw[f][r] = 0.0;
                if (real_col) {
                    const double* src = srow - (long)f * p.pitch;
                    if (VEC) {
#pragma unroll
                        for (int r2 = 0; r2 < R / 2; ++r2) {
                            const double2 v = *reinterpret_cast<const double2*>(src + 2 * r2);
                            fw[f][R - 1 - 2 * r2] = v.x;
                            fw[f][R - 2 - 2 * r2 >= 0 ? R - 2 - 2 * r2 : 0] = v.y;
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < R; ++r) fw[f][r] = src[R - 1 - r];
                    }
                }
            }
        }
        // Cells are visited in ANTI-DIAGONAL order inside the lane's R x F block: the cells of one
        // anti-diagonal are independent (instruction-level parallelism inside one warp) and mostly
        // share (a, -b), so consecutive DFMA/DMUL reuse their coefficient operand.
        double U[R][F];
        double sacc2[RC][F];
#pragma unroll
        for (int rc = 0; rc < RC; ++rc)
#pragma unroll
            for (int f = 0; f < F; ++f) sacc2[rc][f] = 0.0;
#pragma unroll
        for (int dgl = 0; dgl < R + F - 1; ++dgl) {
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const int r = dgl - f;
                if (r >= 0 && r < R) {
                    const int rm = r > 0 ? r - 1 : 0, fm = f > 0 ? f - 1 : 0;
                    const double left = f == 0 ? u[r] : U[r][fm];
                    const double up = r == 0 ? tops[f] : U[rm][f];
                    const double diag = r == 0 ? (f == 0 ? topprev : tops[fm]) : (f == 0 ? u[rm] : U[rm][fm]);
                    double v;
                    if (EXACT) v = cell<true>(left, up, diag, ca[r >> LOGD], cb[r >> LOGD]);
                    else v = fma(ca[r >> LOGD], left + up, cb[r >> LOGD] * diag);   // cb holds -b
                    if (REV) sacc2[r >> LOGD][f] = fma(fw[f][r], diag, sacc2[r >> LOGD][f]);
                    U[r][f] = v;
                    if (MODE == MODE_FWD_STORE) {
                        // u[p, q] = the diagonal input of cell (p, q); stored as soon as it is known
                        double* dst = srow + (long)f * p.pitch;
                        if (VEC) {
                            if (r & 1) {
                                const int r2 = r - 1, r2m = r2 > 0 ? r2 - 1 : 0;
                                const double dprev = r2 == 0 ? (f == 0 ? topprev : tops[fm]) : (f == 0 ? u[r2m] : U[r2m][fm]);
                                if (real_col) *reinterpret_cast<double2*>(dst + r2) = make_double2(dprev, diag);
                            }
                        } else {
                            if (real_col) dst[r] = diag;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) u[r] = U[r][F - 1];
#pragma unroll
        for (int f = 0; f < F; ++f) bots[f] = U[R - 1][f];
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) {
            double t = sacc2[rc][0];
#pragma unroll
            for (int f = 1; f < F; ++f) t += sacc2[rc][f];
            sacc[rc] = t;
        }
        topprev = tops[F - 1];
        if (dummy) {
#pragma unroll
            for (int r = 0; r < R; ++r) u[r] = 1.0;
            topprev = 1.0;
        }

        // ---- 5. outputs ---------------------------------------------------------------------------
        if (BAND) {
            if (p.band_bot != nullptr && svalid && c < N - 1 && lane == tstar) {
                double res = 0.0;
#pragma unroll
                for (int rc = 0; rc < RC; ++rc)
                    if (rc == rcstar) res = u[(rc + 1) * F - 1];
                p.band_bot[(long)sjob * (N - 1) + c] = res;
            }
        }
        if (MODE == MODE_FWD || MODE == MODE_FWD_STORE) {
            if (svalid && c == N - 2 && lane == tstar && (!BAND || p.out != nullptr)) {
                double res = 0.0;
#pragma unroll
                for (int rc = 0; rc < RC; ++rc)
                    if (rc == rcstar) res = u[(rc + 1) * F - 1];
                if (p.pairs == PAIRS_BATCH) {
                    p.out[sa] = res;
                } else {
                    p.out[(long)sa * p.B + sb] = res;
                    if (p.pairs == PAIRS_SYM) p.out[(long)sb * p.B + sa] = res;
                }
            }
        }
        if (REV) {
            // coarse sensitivities of column c (0 on the dummy step and for padded coarse rows)
            double Scur[RC];
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) {
                const bool ok = !dummy && (lane * RC + rc < M - 1);
                Scur[rc] = ok ? sacc[rc] * p.scale4 : 0.0;
            }
            if (MODE == MODE_REV_S) {
                if (real_col) {
                    const long pi = (p.pairs == PAIRS_BATCH) ? (long)sa : (long)sa * p.B + sb;
                    double* Sp = p.S + pi * ((long)(M - 1) * (N - 1));
#pragma unroll
                    for (int rc = 0; rc < RC; ++rc) {
                        const int ip = lane * RC + rc;                   // reversed coarse row
                        if (ip < M - 1) Sp[(long)(M - 2 - ip) * (N - 1) + (N - 2 - c)] = Scur[rc];
                    }
                }
            } else {
                // T[n', c] = S'[n', c] - S'[n'-1, c] - S'[n', c-1] + S'[n'-1, c-1]; W = T * k[n', c]
                const double* yrow = syb + (long)(c < N ? c : N - 1) * Dp;
#pragma unroll
                for (int rc = 0; rc < RC; ++rc) {
                    const double uc = rc > 0 ? Scur[rc > 0 ? rc - 1 : 0] : up_c;
                    const double uc1 = rc > 0 ? Sprev[rc > 0 ? rc - 1 : 0] : up_c1;
                    const double T = (Scur[rc] - uc) - (Sprev[rc] - uc1);
                    const double W = KIND == KIND_RBF ? T * kh3[rc] : T;
                    double* acc = smem + (rc * (D + 1)) * 32 + lane;
                    acc[0] += W;
                    for (int k = 0; k < D; ++k) acc[(k + 1) * 32] = fma(W, __ldg(yrow + 1 + k), acc[(k + 1) * 32]);
                }
                if (c == N - 1) {
                    // the pair is complete for this lane: emit its node rows, clear the accumulators
                    const long pi = (p.pairs == PAIRS_BATCH) ? (long)sa : (long)sa * p.B + sb;
#pragma unroll
                    for (int rc = 0; rc < RC; ++rc) {
                        const int np = lane * RC + rc;                   // reversed node row
                        double* acc = smem + (rc * (D + 1)) * 32 + lane;
                        if (svalid && np < M) {
                            double* gout = p.grad + (pi * M + (M - 1 - np)) * D;
                            const double* xr = sxb + (long)np * Dp;
                            const double sW = acc[0];
                            for (int k = 0; k < D; ++k) {
                                const double gy = acc[(k + 1) * 32];
                                gout[k] = KIND == KIND_RBF ? fma(p.gscale, gy, -(__ldg(xr + 1 + k) * sW)) : p.gscale * gy;
                            }
                        }
                        for (int k = 0; k <= D; ++k) acc[k * 32] = 0.0;
                    }
                }
            }
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) Sprev[rc] = Scur[rc];
            Slast_prev = Slast_cur;
            Slast_cur = Scur[RC - 1];
        }

        // ---- 6. rotate the static-kernel history, advance both streams -----------------------------
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) { kh3[rc] = kh2[rc]; kh2[rc] = kh1[rc]; kh1[rc] = knew[rc]; }
        if (c == NS - 1) {
            // the stencil stream moves on to the pair the production stream is in (it switched
            // exactly once during the last NS steps)
            c = 0;
            sa = a; sb = b; sjob = job; svalid = pvalid && e >= 0;
            syb = yb;
            if (FUSED) sxb = p.Xp + (unsigned)a * xstride;
        } else {
            ++c;
        }
        if (++e == NS) {
            e = 0;
            if (lane == 0) {
                job = job_next;
                pvalid = job < p.njobs;
                if (pvalid) {
                    job_next = (int)(gridDim.x + atomicAdd(p.counter, 1u));
                    job_decode(p, p.job0 + job, a, b);
                }
            } else {
                job = njob; a = na; b = nb_; pvalid = njob < p.njobs;
            }
            if (pvalid) set_ptrs();
        }
    };

#pragma unroll 1
    while (true) {
        // lanes still holding (or draining) a real pair keep the warp alive; the extra steps a trip
        // may run past the end are harmless (every output is guarded by the stream-valid flags)
        const bool alive = pvalid || svalid;
        if (!__any_sync(FULL, alive)) break;
#pragma unroll
        for (int it = 0; it < UNR; ++it) step();
    }
}

}  // namespace skb
