// skb_fwd5.cuh -- the v5 kernel of the fused static kinds (Linear / RBF): the hot path of compute_Gram /
// compute_kernel (MODE 0, every BASELINE config; S1 = _naive_solver as a template flag) and of the adjoint pass behind
// compute_mmd(...).backward(): MODE_FWD_EMIT (forward + last row / column of every grid), MODE_REV_RECON (reversed sweep
// that rebuilds the forward grid backwards) and MODE_REV_RECON_SYM (the same over unordered pairs of Gram(X, X), gradient
// w.r.t. both paths); MODE_FWD_STORE / MODE_REV_GRAD (stored grid) remain as their device-side fallback.
// DESIGN.md 3a / 4a have the measurements behind every choice below.
//
// Same decomposition as solver_kernel (skb_solver.cuh): a warp streams through path pairs, lane t owns RC coarse
// rows (R = RC * 2^d fine rows in registers) and runs one macro step (= one coarse column) behind lane t-1.
// What is different:
//   * lane 0 pops an atomic job queue one pair ahead and publishes (job, byte offsets of X_a, Y_b) in a ring in
//     shared memory; lane t picks its entry up when its own production column wraps: no per-step shuffles of the
//     pair ids, and warps that drift apart are rebalanced (a static assignment lost 24 % to the tail);
//   * the lane's x rows live in registers for the whole pair (when they fit) and the y row of the NEXT
//     production column is loaded one macro step ahead: no load is consumed in the step that issued it;
//   * the static kernel is kept as COLUMN DIFFERENCES d[i][j] = k[i][j+1] - k[i][j], produced pre-scaled by
//     4^-d / sqrt(12); the increment of a coarse cell is one subtraction, its coefficients three more DP
//     instructions, and ONE value (not two) goes up a lane;
//   * the two neighbour exchanges (bottom row down, d of the first row up) go through triple-buffered shared-
//     memory slots written by the PRODUCER as soon as the value exists and read after one warp / block barrier;
//     slot 0 is the boundary u = 1 (no lane-0 select) and a warp boundary is just another slot, which is how
//     2 or 4 warps share one long pair (NW) and how 16 lanes suffice for a short one (LPP = 16: two pair streams
//     per warp, twice the cells per lane and step for the same per-step overhead);
//   * production runs 4 columns ahead of the stencil so that the exchanged d value is one step old;
//   * the three per-pair events (output, boundary re-arm, production wrap) hang off one test per step;
//   * the exp() underflow guard is one unsigned min on the high word; constants are constant-bank operands;
//   * stored-grid adjoint modes: lane-major grid, 256-bit sector stores / loads, the reversed sweep reads its rows
//     through a per-lane cp.async ring 4 steps ahead, gradient accumulators in registers;
//   * reconstruction modes: a second stencil runs the forward recurrence backwards from the emitted boundaries (checked
//     against u[., 0] = 1 per pair); lanes PARK their per-pair sums in shared memory and the warp emits a pair's gradient
//     rows once its last lane is through (the skew makes every per-lane event cost the warp a full pass);
//   * sharded forward: results can go to every rank's copy of G (peer memory), optionally followed by an in-kernel barrier
//     across the ranks (last block signals and waits).
//
// fp64 throughout, FMA arithmetic (u11 = a (u10 + u01) + (-b) u00), results within 1e-13 of solver_kernel.
// Reference semantics replaced: sigkernel/cuda_backend.py:121-160 (+ :6-49), static_kernels.py:17-33, 42-73,
// sigkernel.py:362-364, 607-613 (forward) and :419-502, 256-343 (adjoint) -- see skb_solver.cuh for the mapping.
#pragma once
#include <type_traits>
#include "skb_solver.cuh"

namespace skb {

// (RC, LOGD) shapes fwd5 is instantiated for: the strips that fit 128 registers without spilling (16 resident
// warps per SM); larger strips stay with solver_kernel.  Keep in sync with fwd5_shape_ok() (skb_dispatch.cu).
#define SKB_FWD5_SHAPES(X) X(1, 0) X(1, 1) X(1, 2) X(1, 3) X(2, 0) X(2, 1) X(2, 2) X(4, 0)
// 16-row strips, one warp per pair only (8 resident warps per SM): len_x <= 128 at dyadic order 2 runs faster this
// way than with two warps and a block barrier per step (3.33 vs 3.67 ms on one rank's share of cfg5)
#define SKB_FWD5_R16_SHAPES(X) X(2, 3) X(4, 2) X(8, 1)

// (RC, LOGD) shapes of the 16-lanes-per-pair variant (len_x <= 16 RC): strips of at most 16 fine rows.
// Keep in sync with fwd5_l16_shape_ok() (skb_dispatch.cu).
#define SKB_FWD5_L16_SHAPES(X) X(1, 0) X(1, 1) X(1, 2) X(1, 3) X(2, 0) X(2, 1) X(2, 2) X(2, 3) X(4, 0) X(4, 1) X(4, 2)

// shapes of the adjoint modes: one warp per pair, dyadic order >= 1 (MM and the strips even: 16-byte grid rows).
// Keep in sync with adjoint5_shape_ok() (skb_dispatch.cu).
#define SKB_ADJ5_SHAPES(X) X(1, 1) X(1, 2) X(1, 3) X(2, 1) X(2, 2)

// Table lookup of the two exp variants below.  Entry j of the table holds kscale * 2^(j/TAB) with (j << SH) taken off
// its high word (SH = 20 - log2 TAB), so that adding (ti << SH), ti = TAB n + j, restores the entry AND scales it by
// 2^n in one integer instruction; the entry's shared address is a mask and one multiply-add.
template <int TAB>
__device__ __forceinline__ double exp_tab_entry(int ti, unsigned tab_s) {
    constexpr int SH = TAB == 2048 ? 9 : (TAB == 256 ? 12 : -1);
    static_assert(SH > 0, "table sizes: 256, 2048");
    int lo, hi;
    asm("{\n\t.reg .b32 j, a;\n\tand.b32 j, %2, %3;\n\tmad.lo.u32 a, j, 8, %4;\n\tld.shared.v2.b32 {%0, %1}, [a];\n\t}"
        : "=r"(lo), "=r"(hi) : "r"(ti), "n"(TAB - 1), "r"(tab_s));
    return __hiloint2double(hi + (ti << SH), lo);
}

// exp(x) for x <= ~0, table-driven as exp_neg(); the underflow guard clamps x to >= -700.x through an
// unsigned min on the high word (negative doubles order like their unsigned high words); NaN (canonical,
// sign clear) passes through.
__device__ __forceinline__ double exp_neg5(double x, unsigned tab_s, const KArgs& p) {
    const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
    const unsigned hi = min((unsigned)__double2hiint(x), 0xC085E000u);
    const double xc = __hiloint2double((int)hi, __double2loint(x));
    const double t = fma(xc, p.ek, MAGIC);               // ek = 256 / ln 2
    const double nf = t - MAGIC;
    double r = fma(nf, p.ehi, xc);                       // ehi + elo = -ln2 / 256; n * ehi is exact
    r = fma(nf, p.elo, r);
    double q = fma(r, p.e4, p.e3);                       // e^r - 1 = r (1 + r/2 + r^2/6 + r^3/24)
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = q * r;
    const double tj = exp_tab_entry<EXP_TAB>(__double2loint(t), tab_s);   // lo(t) = 256 n + j: kscale 2^n 2^(j/256)
    return fma(tj, q, tj);
}

// Forward mode (MODE 0): the exp argument arrives pre-scaled by 2048 / ln 2 (folded into the prepared rows), so the
// range reduction is three additions (no hi/lo split of ln 2) and a 2^11-entry table leaves a degree-3 polynomial:
// 7 DP instructions instead of 9.  The guard clamps xs to >= -2031616 (x >= -687.6); NaN passes through.
constexpr int EXP_TAB5 = 2048;
__device__ __forceinline__ double exp_scaled5(double xs, unsigned tab_s, const KArgs& p) {
    const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
    const unsigned hi = min((unsigned)__double2hiint(xs), 0xC13F0000u);
    const double xc = __hiloint2double((int)hi, __double2loint(xs));
    const double t = xc + MAGIC;
    const double nf = t - MAGIC;                         // nearest integer = 2048 n + j
    const double r = xc - nf;                            // exact, |r| <= 1/2
    double q = fma(r, p.e3, p.e4);                       // e4 = c^2/2, e3 = c^3/6, ek = c = ln2 / 2048 (forward mode)
    q = fma(q, r, p.ek);
    q = q * r;                                           // e^(r c) - 1, truncation 3.4e-17
    const double tj = exp_tab_entry<EXP_TAB5>(__double2loint(t), tab_s);
    return fma(tj, q, tj);
}

// shared-memory accesses of the neighbour exchange: 32-bit shared addresses with compile-time offsets, so that
// the per-step address arithmetic disappears (the buffer index is the position in the 3x unrolled loop)
template <int OFF>
__device__ __forceinline__ void sts_f64x2(unsigned base, double a, double b) {
    asm volatile("st.shared.v2.f64 [%0 + %3], {%1, %2};" ::"r"(base), "d"(a), "d"(b), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void lds_f64x2(unsigned base, double& a, double& b) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2 + %3];" : "=d"(a), "=d"(b) : "r"(base), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts_f64(unsigned base, double a) {
    asm volatile("st.shared.f64 [%0 + %2], %1;" ::"r"(base), "d"(a), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ double lds_f64(unsigned base) {
    double a;
    asm volatile("ld.shared.f64 %0, [%1 + %2];" : "=d"(a) : "r"(base), "n"(OFF) : "memory");
    return a;
}

// 256-bit global accesses (sm_100: LDG.256 / STG.256): one whole 32-byte sector per lane and instruction
__device__ __forceinline__ void stg_f64x4(double* ptr, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ldg_f64x4(const double* ptr, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(ptr));
}

// cp.async 16 bytes global -> shared, zero-filled when !valid (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// NW warps (32 NW lanes) share one pair: warp w+1 continues the wavefront of warp w (lane 0 of warp w+1 is
// "lane 32 (w+1)"); the two values that cross the warp boundary every step (bottom row of lane 31 going down,
// d of the next warp's first row going up) go through double-buffered shared memory and one block barrier.
//
// MODE 0            forward only: out[pair] = u[MM, NN]
// MODE_FWD_STORE    additionally stores u[p, q] (the diagonal input of every cell) for the adjoint pass, same
//                   scratch layout as solver_kernel: [job][fine column q][32 R row pitch]
// MODE_REV_GRAD     the same sweep on the REVERSED paths (the reference's flipped-increment solve,
//                   sigkernel.py:438-469), multiplied cell by cell with the stored forward grid (read one step
//                   ahead), reduced to coarse sensitivities S and contracted with the analytic static-kernel
//                   derivative into per-point gradients (sigkernel.py:470-500), see solver_kernel
// MODE_FWD_EMIT     forward that also leaves the LAST ROW and the LAST COLUMN of every pair's grid (KArgs.brow /
//                   .bcol): all the adjoint pass below needs from the forward solve
// MODE_REV_RECON    adjoint pass WITHOUT a stored grid: the reversed sweep carries a second stencil that rebuilds the
//                   forward solution backwards from its last row / column, u00 = (a (u10 + u01) - u11) / b -- the same
//                   3-instruction update with coefficients (a/b, -1/b), same neighbours, same sweep direction as the
//                   reversed-path solve.  Zero grid traffic (the stored-grid pair moves 16 B per fine node through HBM:
//                   4.2 GB per 128 x 128 Gram at len 64, dyadic order 1).  The backward recurrence amplifies rounding
//                   errors by the growth of the solution squared, so every pair checks u[., 0] = 1 at the end of its
//                   sweep and raises KArgs.flag when it misses by more than recon_tol: the host side has already queued
//                   the stored-grid kernels behind a device-side test of that flag (KArgs.cond).  The epilogue can
//                   contract the per-point gradients with d loss / d K on the fly (gradX += coef * grad_points, atomic).
// LPP = lanes per pair: 32 (default), or 16 -- then a warp carries TWO independent pair streams (lanes 0-15 and
// 16-31), each lane owns twice the rows and the per-step overhead is spread over twice the cells.
// the exchange buffer of a step: an std::integral_constant in the 3x unrolled loop (buffer offsets are immediates),
// or a runtime int (UNR == 1: a third of the code, for the modes whose step is too large to unroll)
template <class T> struct step_q {
    static constexpr bool ct = true;
    static constexpr int value = T::value;
    static __device__ __forceinline__ int runtime(T) { return 0; }
};
template <> struct step_q<int> {
    static constexpr bool ct = false;
    static constexpr int value = 0;
    static __device__ __forceinline__ int runtime(int q) { return q; }
};

template <int KIND, int RC, int LOGD, int DP2, int NW, int MINB, int UNR, int MODE = 0, int LPP = 32, bool S1 = false>
__global__ void __launch_bounds__(32 * NW, MINB) fwd5_kernel(const KArgs p) {
    static_assert(!S1 || (MODE == 0 && NW == 1), "scheme S1 (_naive_solver): single-warp forward variants only");
    static_assert(LPP == 32 || (LPP == 16 && NW == 1 && (MODE == 0 || MODE == MODE_FWD_EMIT || MODE == MODE_REV_RECON)),
                  "16 lanes per pair: one warp; forward, forward + boundaries, reconstruction adjoint");
    constexpr int NSTR = 32 / LPP;               // pair streams per warp
    constexpr int F = 1 << LOGD;
    constexpr int R = RC * F;
    constexpr bool STORE = MODE == MODE_FWD_STORE, REVG = MODE == MODE_REV_GRAD;
    // MODE_REV_RECON_SYM: the reversed sweep of Gram(X, X) over the UNORDERED pairs a <= b -- besides d k / d X_a it
    // contracts the same sensitivities with the rows of X_a per node COLUMN (the partial sums travel down the lane chain
    // with the exchange) and so also yields d k / d X_b: one sweep per unordered pair instead of two
    constexpr bool RSYM = MODE == MODE_REV_RECON_SYM;
    constexpr bool EMIT = MODE == MODE_FWD_EMIT, RECON = MODE == MODE_REV_RECON || RSYM;
    constexpr bool REVX = REVG || RECON;          // reversed sweep: sensitivities, gradient epilogue
    static_assert(MODE == 0 || STORE || REVG || EMIT || RECON, "unknown mode");
    static_assert(!(STORE || REVG) || (NW == 1 && LOGD >= 1), "the stored-grid modes use one warp per pair and 16-byte grid rows");
    // REV_GRAD reads the stored grid through a per-lane cp.async ring in shared memory, DEPTH steps ahead (the
    // lane-major layout makes every lane's stream contiguous): the loads of a whole DEPTH-step window are in
    // flight per lane, which is what it takes to cover the loaded HBM latency (measured ~3 us) -- one step ahead
    // in registers was not enough (long_scoreboard 5.2 stalled warps per issue).
    constexpr bool STAGE = REVG && (F * R <= 32);
    constexpr int DEPTH = (F * R <= 8) ? 4 : (F * R <= 16 ? 2 : 1);
    constexpr int NP = F * R / 2;                // 16-byte pieces per lane and step
    constexpr bool GREG = REVX && (RC * DP2 <= 6);   // gradient accumulators in registers instead of shared memory
    // REV_RECON with register accumulators: a lane that finishes a pair only parks its sums (and the rebuilt first column,
    // for the boundary check) in shared memory; the warp emits the gradient rows of the pair together, once its last
    // lane is through -- the epilogue runs once per pair and warp instead of once per pair and LANE (the lanes are
    // skewed, so every per-lane event costs the warp a full pass with one lane active)
    constexpr bool UFLUSH = RECON && GREG;
    static_assert(!RSYM || (UFLUSH && NW == 1 && LPP == 32 && UNR == 3), "the unordered-pair sweep: one warp per pair, register accumulators");
    constexpr int Dp = 2 * DP2;
    constexpr bool XREG = (RC * DP2 <= ((LPP == 16 && R > 8) ? 12 : 8));   // x rows of the pair in registers (the 16-row
                                                                           // strips run 8 warps per SM: room for 12 double2)
    constexpr int LEAD = 4;                      // production column = stencil column + LEAD (mod N)
    constexpr int RING = NW == 1 ? 32 : 64;      // job ring depth (lane 0 is < 32 NW steps = 8 NW wraps < RING/2 ahead; REV_RECON's
                                                 // flush looks LEAD more steps back)
    if ((MODE == MODE_FWD_STORE || MODE == MODE_REV_GRAD) && p.cond != nullptr) {
        // stored-grid passes queued as the fallback of the reconstruction adjoint: run only if it raised its flag
        if (*reinterpret_cast<const volatile unsigned int*>(p.cond) == 0u) return;
    }
    const int glane = threadIdx.x;               // position in the wavefront
    const int lane = glane & 31;
    const int pl = LPP == 32 ? glane : (lane & (LPP - 1));   // position in the pair's wavefront
    const int sid = LPP == 32 ? 0 : lane / LPP;               // pair stream of this lane
    const int slot = LPP == 32 ? glane : sid * (LPP + 1) + pl;   // exchange slot (one boundary slot per stream)
    const int N = p.N, M = p.M;                  // N >= LEAD (the dispatcher sends shorter paths elsewhere)
    const int G = gridDim.x * NSTR;
    const int first_job = blockIdx.x * NSTR + sid;
    const bool has_job = first_job < p.njobs;

    constexpr int Fq = 1 << LOGD, Rq = RC * Fq;
    constexpr bool XREGq = (RC * DP2 <= ((LPP == 16 && Rq > 8) ? 12 : 8));
    // pre-scaled exp argument + 2^11 table: single-warp forward variants with strips of more than 8 rows (8 resident
    // blocks per SM: room for 16 KB each) whose x rows live in registers.  Keep in sync with fwd5_scaled_exp().
    constexpr bool SCALED = (MODE == 0 || EMIT) && NW == 1 && XREGq && Rq > 8;
    constexpr int ETAB = SCALED ? EXP_TAB5 : EXP_TAB;
    __shared__ double etab[ETAB];   // RBF: kscale * 2^(j/ETAB), high word less j << SH (exp_tab_entry)
    const unsigned etab_s = (unsigned)__cvta_generic_to_shared(etab);
    __shared__ int4 ring_s[NSTR][RING];          // job stream: (job, x offset, y offset in bytes, row a of the pair)
    __shared__ longlong2 ring_b[RECON ? NSTR : 1][RECON ? RING : 1];   // REV_RECON: (last row, last column) of the pair's forward grid
    // Neighbour exchange through shared memory, triple-buffered (buffer = position in the 3x unrolled loop;
    // one warp / block barrier per step separates the writes from the reads): lane g writes its bottom row to
    // slot g+1 and reads the row above its strip from slot g -- slot 0 holds the boundary u = 1, so lane 0
    // needs no special case and a warp boundary (NW > 1) is just another slot; likewise the d value of the
    // first node row goes UP one lane.
    static_assert(UNR == 3 || UNR == 1, "the exchange buffers and the d history rotate with period 3 (UNR == 1: runtime buffer index)");
    constexpr int H = (F + 1) / 2;
    constexpr int NL = LPP == 32 ? 32 * NW : NSTR * (LPP + 1) - 1;   // exchange slots - 1
    constexpr int TXH = (NL + 1) * 16, TXQ = H * TXH;     // byte strides of tx[q][h][slot]
    constexpr int DXQ = (NL + 1) * 8;                       // byte stride of dx[q][slot]
    __shared__ double2 tx[3][H][NL + 1];
    __shared__ double dx[3][NL + 1];
    // REV_RECON: the same exchange for the rebuilt forward solution, and the sensitivities of a lane's last coarse row
    // (this and the previous column) going down one lane
    __shared__ double2 txr[RECON ? 3 : 1][RECON ? H : 1][RECON ? NL + 1 : 1];
    __shared__ double2 sxr[RECON ? 3 : 1][RECON ? NL + 1 : 1];
    // REV_RECON_SYM: partial column sums (sum of W, sum of W x_k over the node rows above and including a lane's) going down
    __shared__ double2 syr[RSYM ? 3 : 1][RSYM ? DP2 : 1][RSYM ? NL + 1 : 1];
    // REV_RECON_SYM: the lane's rows of the STENCIL stream's X_a (copied from the production stream's registers when the
    // stencil stream moves on to that pair; re-reading them from global every step stalled on L2: long_scoreboard 1.3 per issue)
    __shared__ double2 xss[RSYM ? RC * DP2 : 1][RSYM ? 32 * NW : 1];
    if (KIND == KIND_RBF) {
        for (int j = glane; j < ETAB; j += 32 * NW) {
            const double e = p.kscale * __ldg(p.exp_tab + j * (2048 / ETAB));
            etab[j] = __hiloint2double(__double2hiint(e) - (j << (ETAB == 2048 ? 9 : 12)), __double2loint(e));
        }
    }

    // ---- job stream -------------------------------------------------------------------------------
    // Lane 0 takes jobs from an atomic queue (one pair ahead, so the atomic's latency is never waited for),
    // decodes (a, b) and publishes (job, offset of X_a, offset of Y_b) in a ring in shared memory; lane t
    // picks entry w up when ITS production column wraps for the w-th time, t steps later (lane 0 is at most
    // 32 NW - 1 steps, i.e. < RING/2 wraps, ahead of the last lane because N >= 4).  Until its first wrap
    // lane t > 0 works on a "virtual" pair (the data of the first real pair, all outputs suppressed).
    const unsigned xstride = (unsigned)(M * Dp * 8), ystride = (unsigned)(N * Dp * 8);   // bytes (< 4 GB: host check)
    int job_next = 0;
    unsigned xo, yo;                              // byte offsets of the production pair's paths
    int pa = 0;                                   // REV_RECON: row a of the production pair (b follows from the job index)
    {
        int a, b;
        job_decode(p, p.job0 + (has_job ? first_job : 0), a, b);
        xo = (unsigned)a * xstride;
        yo = (unsigned)b * ystride;
        pa = a;
        if (pl == 0) {
            for (int q = 0; q < 3; ++q) {
                for (int h = 0; h < H; ++h) tx[q][h][slot] = make_double2(1.0, 1.0);
                dx[q][slot + LPP * (LPP == 32 ? NW : 1)] = 0.0;
                if (RECON) {
                    for (int h = 0; h < H; ++h) txr[RECON ? q : 0][RECON ? h : 0][RECON ? slot : 0] = make_double2(1.0, 1.0);
                    sxr[RECON ? q : 0][RECON ? slot : 0] = make_double2(0.0, 0.0);
                }
                if (RSYM) {
                    for (int i = 0; i < DP2; ++i) syr[RSYM ? q : 0][RSYM ? i : 0][RSYM ? slot : 0] = make_double2(0.0, 0.0);
                }
            }
            ring_s[sid][0] = make_int4(has_job ? first_job : -1, (int)xo, (int)yo, pa);
            job_next = has_job ? (int)(G + atomicAdd(p.counter, 1u)) : p.njobs;
        }
    }
    if (NW > 1) __syncthreads(); else __syncwarp();
    int c = (-pl - LEAD) % N;                     // stencil column; production column e = (c + LEAD) mod N
    if (c < 0) c += N;
    int w = pl == 0 ? 0 : -((pl - 1) / N + 1);    // index of the pair the production stream is in (< 0: virtual)
    int pjob = (pl == 0 && has_job) ? first_job : -1;   // production stream's job (-1: virtual or past the end)
    int sjob = -1;                                // stencil stream's job (-1: nothing to output)
    bool done = pl == 0 && !has_job;
    const int pc = (2 * N - 1 - LEAD) % N;        // stencil column at which the production column wraps
    // this lane holds grid row MM-1 (the output) in u[(orc + 1) * F - 1] iff 0 <= orc < RC
    const int orc = (M - 2) - pl * RC;

    const char* xrow0[RC];                        // this lane's node rows in X_0 (clamped rows never reach a valid cell)
#pragma unroll
    for (int rc = 0; rc < RC; ++rc) {
        int row = pl * RC + rc;
        row = row < M ? row : M - 1;
        xrow0[rc] = reinterpret_cast<const char*>(p.Xp) + (size_t)row * (Dp * 8);
    }
    const unsigned txb0 = (unsigned)__cvta_generic_to_shared(&tx[0][0][slot]);   // read slot; write slot = +16
    const unsigned dxb0 = (unsigned)__cvta_generic_to_shared(&dx[0][slot]);      // write slot; read slot = +8
    const unsigned txrb0 = RECON ? (unsigned)__cvta_generic_to_shared(&txr[0][0][RECON ? slot : 0]) : 0u;
    const unsigned sxrb0 = RECON ? (unsigned)__cvta_generic_to_shared(&sxr[0][RECON ? slot : 0]) : 0u;
    constexpr int SXQ = (NL + 1) * 16;                      // byte stride of sxr[q][slot]
    const unsigned syrb0 = RSYM ? (unsigned)__cvta_generic_to_shared(&syr[0][0][RSYM ? slot : 0]) : 0u;
    constexpr int SYQ = DP2 * (NL + 1) * 16;                // byte stride of syr[q][.][slot]

    double2 xr[XREG ? RC : 1][DP2];
    // !XREG: the lane's rows are staged in shared memory once per pair ([piece][lane]: conflict-free 16-byte
    // accesses) and re-read every step from there -- global loads consumed in the step that issues them were the
    // reason the wide-row shapes (D + 1 = 10) ran latency-bound
    // (else: straight from global / L1 every step; the reconstruction adjoint has its own exchange arrays to fit in)
    constexpr bool XSM = !XREG && (RC * DP2 * 32 * NW * 16 <= (RECON ? 10240 : 20480));
    __shared__ double2 xs_s[XSM ? RC * DP2 : 1][XSM ? 32 * NW : 1];
    const double* xrow[(XREG || XSM) ? 1 : RC];
    const double* yp = p.Yp;                      // y row of the NEXT production column
    auto set_pair = [&]() {
        yp = reinterpret_cast<const double*>(reinterpret_cast<const char*>(p.Yp) + yo);
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) {
            const double* xp = reinterpret_cast<const double*>(xrow0[rc] + xo);
#pragma unroll
            for (int i = 0; i < DP2; ++i) {
                if (XREG) xr[XREG ? rc : 0][i] = ldg2(xp + 2 * i);
                else if (XSM) xs_s[XSM ? rc * DP2 + i : 0][XSM ? glane : 0] = ldg2(xp + 2 * i);
            }
            if (!XREG && !XSM) xrow[(XREG || XSM) ? 0 : rc] = xp;
        }
    };
    set_pair();
    {
        int e = c + LEAD;
        e = e >= N ? e - N : e;
        yp += (unsigned)(e * Dp);
    }
    double2 yq[DP2];
#pragma unroll
    for (int i = 0; i < DP2; ++i) yq[i] = ldg2(yp + 2 * i);

    // ---- adjoint modes -------------------------------------------------------------------------------
    extern __shared__ double gacc[];              // REV_GRAD: accumulators [(rc * (D + 1) + k) * 32 + lane]
    const int D = p.D;
    const long NNf = (long)(N - 1) << LOGD, MMl = (long)(M - 1) << LOGD;
    unsigned sxo = xo, syo = yo;                  // REV_GRAD: byte offsets of the STENCIL stream's paths
    const double* syp = p.Yp;                     // REV_GRAD: y row of the stencil column
    double Sprev[RC], Slast_cur = 0.0, Slast_prev = 0.0;
#pragma unroll
    for (int rc = 0; rc < RC; ++rc) Sprev[rc] = 0.0;
    double fw[REVG ? F : 1][REVG ? R : 1];        // forward values of the cells of this step (reversed row order)
    double ga[GREG ? RC : 1][GREG ? Dp : 1];      // GREG: [coarse row][0: sum of W; 1..D: sum of W y_k]
#pragma unroll
    for (int rc = 0; rc < (GREG ? RC : 1); ++rc)
#pragma unroll
        for (int e2 = 0; e2 < (GREG ? Dp : 1); ++e2) ga[rc][e2] = 0.0;
    constexpr int GL = 32 * NW;                   // lanes of the block: stride of the per-lane shared arrays
    if (REVX) {
        if (!GREG) for (int i = 0; i < RC * (D + 1); ++i) gacc[i * GL + glane] = 0.0;
    }
    if (REVG) {
#pragma unroll
        for (int f = 0; f < F; ++f)
#pragma unroll
            for (int r = 0; r < R; ++r) fw[REVG ? f : 0][REVG ? r : 0] = 0.0;
    }
    // ---- REV_RECON: the rebuilt forward solution ub (reversed coordinates: ub(p', q') = u[MM - p', NN - q']) ----------
    double ub[RECON ? R : 1], topsb[RECON ? F : 1], topprevb = 1.0, bpre[RECON ? F : 1];
    double chk_tol = 0.0;                         // recon_tol * max(1, |k|) of the stencil stream's pair
    double cur_coef = 0.0;                        // fused loss head: d loss / d k of the stencil stream's pair
    double* cur_gx = nullptr;                     //                  and its rows of d loss / d X
#pragma unroll
    for (int r = 0; r < (RECON ? R : 1); ++r) ub[r] = 1.0;
#pragma unroll
    for (int f = 0; f < (RECON ? F : 1); ++f) { topsb[f] = 1.0; bpre[f] = 1.0; }
    // boundary arrays of the stencil stream's pair (cbr: row reached through lane 0's top; set at the re-arm) and of the
    // pair the production stream is in (nbr / nbc: fixed when the production column wraps, 4 steps before the re-arm)
    const double* cbr = p.brow + F;               // lane 0: next step's part of the last row (walks down, F values per step)
    bool cbr_ok = false;
    const double* nbr = p.brow;
    const double* nbc = p.bcol;
    // staging of a lane's first column u[., NN] (R + 1 values incl. the node above the strip, then u[MM, NN] = k itself,
    // the scale of the boundary check): [k][lane] behind gacc
    double* const bstg = RECON ? gacc + (GREG ? 0 : (size_t)RC * (D + 1) * GL) : nullptr;
    // UFLUSH: parked sums [buffer][rc * DP2 + i][lane] (double2) and first columns [buffer][r][lane], buffer = pair index
    // mod (fbuf_mask + 1) -- as many buffers as pairs fit between a lane's event and the flush of its warp (host: N (mask + 1) >= 34)
    constexpr int FBUF = (RC * Dp + R) * GL;      // doubles per buffer
    double* const gst = UFLUSH ? bstg + (size_t)(R + 2) * GL : nullptr;
    // REV_RECON_SYM: the completed column sums of the pair the last lane is in, [node column][Dp], behind the parked sums
    double* const ycol = RSYM ? gst + (size_t)(p.fbuf_mask + 1) * FBUF : nullptr;
    const unsigned ycol_s = RSYM ? (unsigned)__cvta_generic_to_shared(ycol) : 0u;
    auto pair_boundaries = [&](int job_) {
        // slot of the pair in the forward launch's boundary arrays; under bsym the pair (a, b), a > b, reads the
        // transposed grid of (b, a)
        if (RECON) {
            long sl = job_ >= 0 ? p.job0 + job_ : 0;      // (virtual / past the end: slot 0, nothing is read from it)
            bool swp = false;
            if (p.bsym && job_ >= 0) {
                const int a = pa, b = (int)(sl - (long)pa * p.B);      // bsym: GRAM enumeration (no division: a rides in the ring)
                swp = a > b;
                const long lo = swp ? b : a, hi = swp ? a : b;
                sl = lo * p.A - lo * (lo - 1) / 2 + (hi - lo);
            }
            const double* br = p.brow + sl * p.brow_stride;
            const double* bc = p.bcol + sl * p.bcol_stride;
            nbr = swp ? bc : br;
            nbc = swp ? br : bc;
        }
    };
    auto stage_first_column = [&](bool real) {
        if (RECON) {
            // ub[r] = u[MM - pl R - r - 1, NN] (r = -1 .. R-1) = nbc[MM - pl R - R + k], k = 0 .. R
            // entries of rows above the grid (strips past its end) are zero-filled: no bytes are read for them, and the
            // clamped base keeps even their addresses inside the boundary arrays
            const long i0 = MMl - (long)(pl + 1) * R;
            const double* nb = nbc + (i0 < -(R + 1) ? -(R + 1) : (int)i0);
            const int kmin = i0 < -(R + 1) ? R + 1 : (i0 < 0 ? (int)-i0 : 0);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(bstg + glane);
#pragma unroll
            for (int k = 0; k <= (UFLUSH ? R : R + 1); ++k) {
                const bool ok = real && (k > R || k >= kmin);
                const double* src = k > R ? nbc + MMl : nb + k;
                const int n8 = ok ? 8 : 0;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + (unsigned)(k * GL * 8)), "l"(src), "r"(n8) : "memory");
            }
            cp_async_commit();
        }
    };
    if (RECON) {
        // lane 0 of a pair starts inside its first pair (no production wrap before the first re-arm)
        pair_boundaries(pjob);
        if (pl == 0) ring_b[RECON ? sid : 0][0] = make_longlong2((long long)nbr, (long long)nbc);   // read >= 1 step (1 barrier) later
        stage_first_column(pjob >= 0);
    }
    // Layout of the stored grid (v5 adjoint): LANE-major, [job][forward lane t][fine column q][R rows] -- each
    // lane streams through its own contiguous NNf * R doubles, forwards when storing, backwards when the
    // reversed sweep reads them.  (The lanes of a warp sit at 32 different columns, so nothing coalesces across
    // lanes anyway; the column-major layout of solver_kernel makes every lane touch a different 1 KB row per
    // step -- one DRAM page activation per 32-64 bytes.  Measured at cfg4: see DESIGN.md.)
    // The reversed lane t reads forward rows p0 .. p0 + R - 1, p0 = MMl - (t+1) R: they straddle two forward
    // lanes when MMl is not a multiple of R; rows p0 < 0 (strip outside the grid) are skipped, their S is masked.
    const long lane_stride = NNf * R, job_stride = 32 * lane_stride;
    auto load_fw = [&](int job_, int col_) {
        if (REVG) {
            const bool real = job_ >= 0 && col_ < N - 1;
            const double* jb = p.scratch + (long)(real ? job_ : 0) * job_stride + (NNf - 1 - (long)(real ? col_ : 0) * F) * R;
#pragma unroll
            for (int f = 0; f < F; ++f) {
                if (R % 4 == 0 && (MMl & 3) == 0) {
                    // MMl a multiple of 4: the reversed strips line up with the 32-byte groups of the forward lanes
#pragma unroll
                    for (int j = 0; j < R / 4; ++j) {
                        const long p0 = MMl - (long)(lane + 1) * R + 4 * j;
                        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
                        if (real && p0 >= 0)
                            ldg_f64x4(jb + (p0 / R) * lane_stride - (long)f * R + (p0 % R), v0, v1, v2, v3);
                        // forward rows p0 .. p0+3 are reversed rows R-1-4j .. R-4-4j of this lane's strip
                        fw[REVG ? f : 0][REVG ? R - 1 - 4 * j : 0] = v0;
                        fw[REVG ? f : 0][REVG ? (R - 2 - 4 * j >= 0 ? R - 2 - 4 * j : 0) : 0] = v1;
                        fw[REVG ? f : 0][REVG ? (R - 3 - 4 * j >= 0 ? R - 3 - 4 * j : 0) : 0] = v2;
                        fw[REVG ? f : 0][REVG ? (R - 4 - 4 * j >= 0 ? R - 4 - 4 * j : 0) : 0] = v3;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < R / 2; ++j) {
                        const long p0 = MMl - (long)(lane + 1) * R + 2 * j;
                        double2 v = make_double2(0.0, 0.0);
                        if (real && p0 >= 0)
                            v = *reinterpret_cast<const double2*>(jb + (p0 / R) * lane_stride - (long)f * R + (p0 % R));
                        // forward rows p0, p0 + 1 are reversed rows R-1-2j, R-2-2j of this lane's strip
                        fw[REVG ? f : 0][REVG ? R - 1 - 2 * j : 0] = v.x;
                        fw[REVG ? f : 0][REVG ? (R - 2 - 2 * j >= 0 ? R - 2 - 2 * j : 0) : 0] = v.y;
                    }
                }
            }
        }
    };

    // staging ring: [slot][piece k = f * R/2 + j][lane] 16-byte entries, behind the gradient accumulators
    const unsigned stg0 = STAGE ? (unsigned)__cvta_generic_to_shared(gacc + (size_t)RC * (D + 1) * 32) + lane * 16 : 0;
    int sq = 0;                                   // ring slot of the current step
    auto stage_issue = [&](int slot, int job_, int col_) {
        if (STAGE) {
            const bool real = job_ >= 0 && col_ >= 0 && col_ < N - 1;
            const double* jb = p.scratch + (long)(real ? job_ : 0) * job_stride + (NNf - 1 - (long)(real ? col_ : 0) * F) * R;
#pragma unroll
            for (int f = 0; f < F; ++f) {
#pragma unroll
                for (int j = 0; j < R / 2; ++j) {
                    const long p0 = MMl - (long)(lane + 1) * R + 2 * j;
                    const bool ok = real && p0 >= 0;
                    const double* src = ok ? jb + (p0 / R) * lane_stride - (long)f * R + (p0 % R) : p.scratch;
                    cp_async16(stg0 + ((slot * NP + f * (R / 2) + j) * 32) * 16, src, ok);
                }
            }
            cp_async_commit();
        }
    };
    auto stage_consume = [&](int slot) {
        if (STAGE) {
            cp_async_wait<DEPTH - 1>();
#pragma unroll
            for (int f = 0; f < F; ++f) {
#pragma unroll
                for (int j = 0; j < R / 2; ++j) {
                    double vx, vy;
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(stg0 + ((slot * NP + f * (R / 2) + j) * 32) * 16) : "memory");
                    // forward rows p0, p0 + 1 are reversed rows R-1-2j, R-2-2j of this lane's strip
                    fw[REVG ? f : 0][REVG ? R - 1 - 2 * j : 0] = vx;
                    fw[REVG ? f : 0][REVG ? (R - 2 - 2 * j >= 0 ? R - 2 - 2 * j : 0) : 0] = vy;
                }
            }
        }
    };
    if (STAGE) {
#pragma unroll
        for (int d0 = 0; d0 < DEPTH; ++d0) stage_issue(d0, -1, 0);     // the first DEPTH steps are virtual for every lane
    }

    double* ebr = p.brow;                         // FWD_EMIT: where this step's part of the pair's last row goes
    double u[R];
#pragma unroll
    for (int r = 0; r < R; ++r) u[r] = 1.0;
    double tops[F];                               // row above the lane's strip, this step's fine columns
#pragma unroll
    for (int f = 0; f < F; ++f) tops[f] = 1.0;
    double topprev = 1.0;
    // static kernel, pre-scaled by kscale = 4^-d / sqrt(12): k at the newest column; column differences
    // d[j] = k[j+1] - k[j] at the stencil columns c, c+1, c+2 (REV_GRAD: k itself at those columns, because the
    // gradient epilogue needs k[., c] bit-for-bit independent of the neighbouring pairs of the stream)
    double klast[RC], dA[RC], dB[RC], dC[RC];
#pragma unroll
    for (int rc = 0; rc < RC; ++rc) klast[rc] = dA[rc] = dB[rc] = dC[rc] = 0.0;
    double dn = 0.0;                              // d[c] of lane+1's first row

    // (NW > 1: split arrive/sync named barriers were measured slower than the plain block barrier: 6.1 vs
    // 4.65 ms at 64x512 pairs of len 128.)
    auto step = [&](auto qc) __attribute__((always_inline)) {
        using QT = step_q<decltype(qc)>;
        constexpr int Q = QT::value;              // exchange buffer of this step (0 + a runtime part when UNR == 1)
        const int qr = QT::runtime(qc);
        const unsigned txb = txb0 + (unsigned)(qr * TXQ), dxb = dxb0 + (unsigned)(qr * DXQ);
        const unsigned txrb = txrb0 + (unsigned)(qr * TXQ), sxrb = sxrb0 + (unsigned)(qr * SXQ);
        // next step's stencil column is c+1: lane-1 needs this lane's first-row d[c+1] = dC as of NOW (made
        // one step ago), so this exchange does not wait for this step's production
        sts_f64<Q * DXQ>(dxb, REVX ? klast[0] - dC[0] : dC[0]);
        if (RECON) {
            // lane 0 of the pair: the rebuilt row above its strip is the forward solution's LAST ROW -- the values of
            // the NEXT step are loaded now (a step of latency hiding) and replace what the exchange delivers.  cbr walks
            // down the row, F values per step; on the step without a coarse column it jumps to the next pair's row.
            // (cbr jumps to the next pair's row in the event block of the step before the column without a coarse column)
            const bool okb = pl == 0 && cbr_ok && c != N - 2;
#pragma unroll
            for (int f = 0; f < (RECON ? F : 1); ++f) bpre[f] = okb ? __ldg(cbr - f) : 1.0;
            cbr -= F;
        }
        const bool real_col = sjob >= 0 && c < N - 1;
        if (REVG && !STAGE) load_fw(sjob, c);
        if (STAGE) stage_consume(sq);
        double kc[RC];                            // REV_GRAD: k at (own node rows, node column c)
        double up_c = 0.0, up_c1 = 0.0;           // REV_GRAD: S of lane-1's last coarse row at columns c, c-1
        double sacc2[REVX ? RC : 1][REVX ? F : 1];
        if (REVX) {
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) {
                kc[rc] = dA[rc];                  // the reversed sweeps keep the k history itself (dA, dB, dC = k at columns c, c+1, c+2)
#pragma unroll
                for (int f = 0; f < F; ++f) sacc2[REVX ? rc : 0][REVX ? f : 0] = 0.0;
            }
        }
        if (REVG) {
            up_c = shfl_up1(Slast_cur);
            up_c1 = shfl_up1(Slast_prev);
            if (lane == 0) { up_c = 0.0; up_c1 = 0.0; }
        }
        double* srow = nullptr;                   // STORE: this lane's R values of fine column c * F
        if (STORE) srow = p.scratch + (long)(real_col ? sjob : 0) * job_stride + (long)lane * lane_stride + (long)(real_col ? c : 0) * F * R;
        // ---- 1. stencil coefficients of coarse column c ---------------------------------------------
        // e = g / sqrt(12) (g = the refined increment):  -b = e^2 - 1,  a = 1 + g/2 + g^2/12 = sqrt(3) e + (2 - b)
        double ca[RC], cb[RC];
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) {
            // REV_GRAD: the history holds k, so d[c] = k[c+1] - k[c] is formed here (same operands, same rounding
            // as in the other modes, where it is formed once at production time)
            const double dlo = REVX ? dB[rc] - dA[rc] : dA[rc];
            const double dhi = rc + 1 < RC ? (REVX ? dB[rc + 1 < RC ? rc + 1 : rc] - dA[rc + 1 < RC ? rc + 1 : rc] : dA[rc + 1 < RC ? rc + 1 : rc]) : dn;
            const double el = dhi - dlo;
            if (S1) {
                // _naive_solver (cython_backend.pyx:27): a = 1 + g/2, b = 1
                cb[rc] = -1.0;
                ca[rc] = fma(el, p.sqrt3, 1.0);
            } else {
                cb[rc] = fma(el, el, -1.0);
                ca[rc] = fma(el, p.sqrt3, cb[rc] + 2.0);
            }
        }
        // REV_RECON: coefficients of the backward update u00 = (a/b)(u10 + u01) - (1/b) u11:  cib = -1/b = 1/cb, cia = -ca cib
        double cia[RECON ? RC : 1], cib[RECON ? RC : 1];
        if (RECON) {
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) {
                double r0;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(cb[rc]));
                double e0 = fma(-cb[rc], r0, 1.0);
                r0 = fma(r0, e0, r0);
                e0 = fma(-cb[rc], r0, 1.0);
                r0 = fma(r0, e0, r0);
                cib[RECON ? rc : 0] = r0;
                cia[RECON ? rc : 0] = -(ca[rc] * r0);
            }
        }

        // ---- 2. the stencil: R rows x F fine columns in registers, anti-diagonal order ---------------
        double U[R][F];
        double UB[RECON ? R : 1][RECON ? F : 1];
#pragma unroll
        for (int dgl = 0; dgl < R + F - 1; ++dgl) {
            double ss[F], tt[F];
            double ssb[RECON ? F : 1], ttb[RECON ? F : 1];
            if (RECON) {
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    const int r = dgl - f;
                    if (r >= 0 && r < R) {
                        const int rm = r > 0 ? r - 1 : 0, fm = f > 0 ? f - 1 : 0;
                        const double leftb = f == 0 ? ub[RECON ? r : 0] : UB[RECON ? r : 0][RECON ? fm : 0];
                        const double upb = r == 0 ? topsb[RECON ? f : 0] : UB[RECON ? rm : 0][RECON ? f : 0];
                        const double diagb = r == 0 ? (f == 0 ? topprevb : topsb[RECON ? fm : 0])
                                                    : (f == 0 ? ub[RECON ? rm : 0] : UB[RECON ? rm : 0][RECON ? fm : 0]);
                        ssb[RECON ? f : 0] = leftb + upb;
                        ttb[RECON ? f : 0] = cib[RECON ? (r >> LOGD) : 0] * diagb;
                    }
                }
            }
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const int r = dgl - f;
                if (r >= 0 && r < R) {
                    const int rm = r > 0 ? r - 1 : 0, fm = f > 0 ? f - 1 : 0;
                    const double left = f == 0 ? u[r] : U[r][fm];
                    const double up = r == 0 ? tops[f] : U[rm][f];
                    ss[f] = (f & 1) ? up + left : left + up;
                }
            }
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const int r = dgl - f;
                if (r >= 0 && r < R) {
                    const int rm = r > 0 ? r - 1 : 0, fm = f > 0 ? f - 1 : 0;
                    const double diag = r == 0 ? (f == 0 ? topprev : tops[fm]) : (f == 0 ? u[rm] : U[rm][fm]);
                    tt[f] = cb[r >> LOGD] * diag;
                    if (REVG) sacc2[REVG ? (r >> LOGD) : 0][REVG ? f : 0] = fma(fw[REVG ? f : 0][REVG ? r : 0], diag, sacc2[REVG ? (r >> LOGD) : 0][REVG ? f : 0]);
                    if (RECON) {
                        // the rebuilt forward value of this cell's far corner times the reversed solution at its near corner
                        const double ubn = fma(cia[RECON ? (r >> LOGD) : 0], ssb[RECON ? f : 0], ttb[RECON ? f : 0]);
                        UB[RECON ? r : 0][RECON ? f : 0] = ubn;
                        sacc2[RECON ? (r >> LOGD) : 0][RECON ? f : 0] = fma(ubn, diag, sacc2[RECON ? (r >> LOGD) : 0][RECON ? f : 0]);
                        if (r == R - 1 && ((f & 1) || f == F - 1)) {
                            const int f0 = f & ~1;
                            const double v0 = UB[RECON ? r : 0][RECON ? f0 : 0], v1 = UB[RECON ? r : 0][RECON ? f : 0];
                            if (f0 == 0) sts_f64x2<Q * TXQ + 16>(txrb, v0, v1);
                            if (f0 == 2) sts_f64x2<Q * TXQ + TXH + 16>(txrb, v0, v1);
                            if (f0 == 4) sts_f64x2<Q * TXQ + 2 * TXH + 16>(txrb, v0, v1);
                            if (f0 == 6) sts_f64x2<Q * TXQ + 3 * TXH + 16>(txrb, v0, v1);
                        }
                    }
                    if (STORE && (R % 4 == 0 ? (r & 3) == 3 : (r & 1))) {
                        // u[p, q] = the diagonal input of cell (p, q); a lane's rows of one fine column leave as whole
                        // 32-byte sectors (R % 4 == 0) or 16-byte pairs, as soon as the last of them is known
                        auto dg = [&](int rr) {
                            const int rrm = rr > 0 ? rr - 1 : 0;
                            return rr == 0 ? (f == 0 ? topprev : tops[fm]) : (f == 0 ? u[rrm] : U[rrm][fm]);
                        };
                        if (real_col) {
                            if (R % 4 == 0) stg_f64x4(srow + f * R + (r - 3), dg(r - 3 >= 0 ? r - 3 : 0), dg(r - 2 >= 0 ? r - 2 : 0), dg(r - 1), diag);
                            else *reinterpret_cast<double2*>(srow + f * R + (r - 1)) = make_double2(dg(r - 1), diag);
                        }
                    }
                }
            }
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const int r = dgl - f;
                if (r >= 0 && r < R) {
                    // (measured: a DFMA with three distinct register operands costs one issue cycle more
                    //  than a two-operand DP instruction; DMUL + DADD instead of it is slower still)
                    U[r][f] = fma(ca[r >> LOGD], ss[f], tt[f]);
                    if (r == R - 1 && ((f & 1) || f == F - 1)) {
                        // hand the bottom-row values to lane+1 as soon as a pair of them exists
                        if (f == 0) sts_f64x2<Q * TXQ + 16>(txb, U[r][0], U[r][0]);
                        if (f == 1) sts_f64x2<Q * TXQ + 16>(txb, U[r][0], U[r][1]);
                        if (f == 2) sts_f64x2<Q * TXQ + TXH + 16>(txb, U[r][2], U[r][2]);
                        if (f == 3) sts_f64x2<Q * TXQ + TXH + 16>(txb, U[r][2], U[r][3]);
                        if (f == 4) sts_f64x2<Q * TXQ + 2 * TXH + 16>(txb, U[r][4], U[r][4]);
                        if (f == 5) sts_f64x2<Q * TXQ + 2 * TXH + 16>(txb, U[r][4], U[r][5]);
                        if (f == 6) sts_f64x2<Q * TXQ + 3 * TXH + 16>(txb, U[r][6], U[r][6]);
                        if (f == 7) sts_f64x2<Q * TXQ + 3 * TXH + 16>(txb, U[r][6], U[r][7]);
                    }
                }
            }
        }
        topprev = tops[F - 1];
#pragma unroll
        for (int r = 0; r < R; ++r) u[r] = U[r][F - 1];
        if (RECON) {
            topprevb = topsb[RECON ? F - 1 : 0];
#pragma unroll
            for (int r = 0; r < R; ++r) ub[RECON ? r : 0] = UB[RECON ? r : 0][RECON ? F - 1 : 0];
        }
        if (EMIT) {
            // the lane that owns grid row MM-1 leaves u[MM, c F + 1 ..] (the last row of the grid) behind
            // (ebr walks along the pair's row, F values per step; it is re-aimed when the stencil stream moves on to a pair, and
            //  u[MM, 0] = 1 is written with the last column)
            if ((unsigned)orc < (unsigned)RC && real_col) {
#pragma unroll
                for (int rc = 0; rc < RC; ++rc)
                    if (rc == orc) {
#pragma unroll
                        for (int f = 0; f < F; ++f) ebr[f] = U[(rc + 1) * F - 1][f];
                    }
            }
            ebr += F;
        }
        if (NW > 1) __syncthreads(); else __syncwarp();
        dn = lds_f64<Q * DXQ + 8>(dxb);
        if (RECON) {
            double vx, vy;
            lds_f64x2<Q * TXQ>(txrb, vx, vy);
            topsb[0] = vx;
            if (F > 1) topsb[RECON && F > 1 ? 1 : 0] = vy;
            if (F > 2) { lds_f64x2<Q * TXQ + TXH>(txrb, vx, vy); topsb[RECON && F > 2 ? 2 : 0] = vx; topsb[RECON && F > 3 ? 3 : 0] = vy; }
            if (F > 4) {
                lds_f64x2<Q * TXQ + 2 * TXH>(txrb, vx, vy); topsb[RECON && F > 4 ? 4 : 0] = vx; topsb[RECON && F > 5 ? 5 : 0] = vy;
                lds_f64x2<Q * TXQ + 3 * TXH>(txrb, vx, vy); topsb[RECON && F > 6 ? 6 : 0] = vx; topsb[RECON && F > 7 ? 7 : 0] = vy;
            }
            if (pl == 0) {
#pragma unroll
                for (int f = 0; f < (RECON ? F : 1); ++f) topsb[f] = bpre[f];
            }
            // sensitivities of lane-1's last coarse row at the columns it finished in ITS previous step (= this lane's
            // columns c and c-1): written after the barrier of that step, read after this one
            if (QT::ct) lds_f64x2<((Q + 2) % 3) * SXQ>(sxrb0, up_c, up_c1);
            else lds_f64x2<0>(sxrb0 + (unsigned)((qr == 0 ? 2 : qr - 1) * SXQ), up_c, up_c1);
        }
        double ycs[RSYM ? Dp : 1];                // REV_RECON_SYM: column sums over the node rows above this lane's (column c)
        if (RSYM) {
#pragma unroll
            for (int i = 0; i < (RSYM ? DP2 : 0); ++i) {
                double vx, vy;
                if (i == 0) lds_f64x2<((Q + 2) % 3) * SYQ>(syrb0, vx, vy);
                if (i == 1) lds_f64x2<((Q + 2) % 3) * SYQ + TXH>(syrb0, vx, vy);
                if (i == 2) lds_f64x2<((Q + 2) % 3) * SYQ + 2 * TXH>(syrb0, vx, vy);
                if (i == 3) lds_f64x2<((Q + 2) % 3) * SYQ + 3 * TXH>(syrb0, vx, vy);
                if (i == 4) lds_f64x2<((Q + 2) % 3) * SYQ + 4 * TXH>(syrb0, vx, vy);
                ycs[RSYM ? 2 * i : 0] = vx;        // (slot 0, what lane 0 reads, is zero for good)
                ycs[RSYM ? 2 * i + 1 : 0] = vy;
            }
        }
        {
            double vx, vy;
            lds_f64x2<Q * TXQ>(txb, vx, vy);
            tops[0] = vx;
            if (F > 1) tops[F > 1 ? 1 : 0] = vy;
            if (F > 2) { lds_f64x2<Q * TXQ + TXH>(txb, vx, vy); tops[F > 2 ? 2 : 0] = vx; tops[F > 3 ? 3 : 0] = vy; }
            if (F > 4) {
                lds_f64x2<Q * TXQ + 2 * TXH>(txb, vx, vy); tops[F > 4 ? 4 : 0] = vx; tops[F > 5 ? 5 : 0] = vy;
                lds_f64x2<Q * TXQ + 3 * TXH>(txb, vx, vy); tops[F > 6 ? 6 : 0] = vx; tops[F > 7 ? 7 : 0] = vy;
            }
        }

        // ---- 2b. REV_GRAD: coarse sensitivities of column c -> second difference T -> W = T k -> accumulate ----
        if (REVX) {
            const bool dummy = c >= N - 1;
            double Scur[RC];
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) {
                double t = sacc2[REVX ? rc : 0][0];
#pragma unroll
                for (int f = 1; f < F; ++f) t += sacc2[REVX ? rc : 0][REVX ? f : 0];
                const bool ok = !dummy && (pl * RC + rc < M - 1);
                Scur[rc] = ok ? t * p.scale4 : 0.0;
            }
            double2 ysv[GREG ? DP2 : 1];
            if (GREG) {
#pragma unroll
                for (int i = 0; i < DP2; ++i) ysv[GREG ? i : 0] = ldg2(syp + 2 * i);
            }
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) {
                const double uc = rc > 0 ? Scur[rc > 0 ? rc - 1 : 0] : up_c;
                const double uc1 = rc > 0 ? Sprev[rc > 0 ? rc - 1 : 0] : up_c1;
                const double T = (Scur[rc] - uc) - (Sprev[rc] - uc1);
                const double W = KIND == KIND_RBF ? T * kc[rc] : T;
                if (RSYM) {
                    // the same W against this lane's rows of X_a (the STENCIL stream's pair: the register copy belongs to the
                    // production stream): column sums for d k / d X_b at node column c
#pragma unroll
                    for (int i = 0; i < DP2; ++i) {
                        const double2 xv = xss[RSYM ? rc * DP2 + i : 0][RSYM ? glane : 0];
                        ycs[RSYM ? 2 * i : 0] = fma(W, i == 0 ? 1.0 : xv.x, ycs[RSYM ? 2 * i : 0]);
                        ycs[RSYM ? 2 * i + 1 : 0] = fma(W, xv.y, ycs[RSYM ? 2 * i + 1 : 0]);
                    }
                }
                if (GREG) {
                    // the prepared y row is (norm term, y_1 .. y_D, 0 ...): slot 0 accumulates W itself
#pragma unroll
                    for (int i = 0; i < DP2; ++i) {
                        const double2 yv = ysv[GREG ? i : 0];
                        ga[GREG ? rc : 0][GREG ? 2 * i : 0] = fma(W, i == 0 ? 1.0 : yv.x, ga[GREG ? rc : 0][GREG ? 2 * i : 0]);
                        ga[GREG ? rc : 0][GREG ? 2 * i + 1 : 0] = fma(W, yv.y, ga[GREG ? rc : 0][GREG ? 2 * i + 1 : 0]);
                    }
                } else {
                    double* acc = gacc + (rc * (D + 1)) * GL + glane;
                    acc[0] += W;
                    for (int k = 0; k < D; ++k) acc[(k + 1) * GL] = fma(W, __ldg(syp + 1 + k), acc[(k + 1) * GL]);
                }
            }
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) Sprev[rc] = Scur[rc];
            Slast_prev = Slast_cur;
            Slast_cur = Scur[RC - 1];
            if (RECON) sts_f64x2<Q * SXQ + 16>(sxrb, Slast_cur, Slast_prev);
            if (RSYM) {
                // the same W, contracted with this lane's rows of X_a: column sums for d k / d X_b at node column c
                const unsigned syrb = syrb0 + (unsigned)(qr * SYQ);
#pragma unroll
                for (int i = 0; i < DP2; ++i) {
                    if (i == 0) sts_f64x2<Q * SYQ + 16>(syrb, ycs[0], ycs[RSYM ? 1 : 0]);
                    if (i == 1) sts_f64x2<Q * SYQ + TXH + 16>(syrb, ycs[RSYM ? 2 : 0], ycs[RSYM ? 3 : 0]);
                    if (i == 2) sts_f64x2<Q * SYQ + 2 * TXH + 16>(syrb, ycs[RSYM && Dp > 4 ? 4 : 0], ycs[RSYM && Dp > 4 ? 5 : 0]);
                    if (i == 3) sts_f64x2<Q * SYQ + 3 * TXH + 16>(syrb, ycs[RSYM && Dp > 6 ? 6 : 0], ycs[RSYM && Dp > 6 ? 7 : 0]);
                    if (i == 4) sts_f64x2<Q * SYQ + 4 * TXH + 16>(syrb, ycs[RSYM && Dp > 8 ? 8 : 0], ycs[RSYM && Dp > 8 ? 9 : 0]);
                }
                if (pl == LPP - 1) {
                    const unsigned yc = ycol_s + (unsigned)(c * (Dp * 8));
#pragma unroll
                    for (int i = 0; i < DP2; ++i)
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(yc + 16 * i), "d"(ycs[RSYM ? 2 * i : 0]), "d"(ycs[RSYM ? 2 * i + 1 : 0]) : "memory");
                }
            }
            syp += Dp;
        }

        // ---- 3. production: static kernel at node column e = c + LEAD (y row loaded one step ago) -----
        double dnew[RC];
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) {
            double2 xv = XREG ? xr[XREG ? rc : 0][0] : (XSM ? xs_s[XSM ? rc * DP2 : 0][XSM ? glane : 0] : ldg2(xrow[(XREG || XSM) ? 0 : rc]));
            double acc = fma(xv.y, yq[0].y, xv.x + yq[0].x);
#pragma unroll
            for (int i = 1; i < DP2; ++i) {
                xv = XREG ? xr[XREG ? rc : 0][i] : (XSM ? xs_s[XSM ? rc * DP2 + i : 0][XSM ? glane : 0] : ldg2(xrow[(XREG || XSM) ? 0 : rc] + 2 * i));
                acc = fma(xv.y, yq[i].y, fma(xv.x, yq[i].x, acc));
            }
            if (KIND == KIND_RBF) acc = SCALED ? exp_scaled5(acc, etab_s, p) : exp_neg5(acc, etab_s, p);
            dnew[rc] = REVX ? klast[rc] : acc - klast[rc];      // reversed sweeps: the history rotates k itself
            klast[rc] = acc;
        }

        // ---- 4. rotate the d history ---------------------------------------------------------------------
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) { dA[rc] = dB[rc]; dB[rc] = dC[rc]; dC[rc] = dnew[rc]; }

        // ---- 5. advance; the rare events of a pair all hang off one test -------------------------------
        // (a lane sees each event once per pair, but the lanes are skewed: the warp runs each block below
        //  in about half of its steps, for one lane at a time -- they are kept as short as possible)
        const int cc = c;
        ++c;
        yp += Dp;
        if (cc >= N - 2 || cc == pc) {
            if (EMIT && cc == N - 2 && sjob >= 0) {
                // last coarse column done: this lane's part of the LAST COLUMN u[., NN] of the grid
                double* bc = p.bcol + (p.job0 + sjob) * p.bcol_stride + 1 + (long)pl * R;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if ((long)pl * R + r < MMl) bc[r] = u[r];
                if (pl == 0) bc[-1] = 1.0;
                if ((unsigned)orc < (unsigned)RC) p.brow[(p.job0 + sjob) * p.brow_stride] = 1.0;
            }
            if (RECON && cc == N - 2 && pl == 0) {
                cbr = nbr + (NNf - 1);
                cbr_ok = pjob >= 0;
            }
            if (UFLUSH && cc == N - 2) {
                // the rebuilt first column u[., 0] of the pair (checked against the boundary value 1 in the flush)
                double* us = gst + (size_t)((w - 1) & p.fbuf_mask) * FBUF + RC * Dp * GL + glane;
#pragma unroll
                for (int r = 0; r < R; ++r) us[r * GL] = ub[RECON ? r : 0];
            }
            if (RECON && !UFLUSH && cc == N - 2 && sjob >= 0) {
                // the rebuilt grid must end at the boundary u[., 0] = 1: a miss beyond recon_tol * max(1, |k|) (errors scale with
                // the size of the solution; or a NaN) sends the whole call to the stored-grid kernels queued behind this one
                bool bad = false;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if ((long)pl * R + r < MMl) bad = bad || !(fabs(ub[RECON ? r : 0] - 1.0) <= chk_tol);
                // (a pair that contributes nothing -- zero weight in the loss head, no per-point gradients asked for -- may
                //  miss: the diagonal of Gram(X, X) grows by orders of magnitude more than the rest and carries no weight
                //  in the MMD and the scoring rules)
                if (bad && (p.grad != nullptr || p.gradX == nullptr || cur_coef != 0.0)) *p.flag = 1u;
            }
            if (!REVX && cc == N - 2) {
                // last coarse column done: u[MM, NN] is in the lane that owns grid row MM-1
                if (sjob >= 0 && (unsigned)orc < (unsigned)RC) {
                    double res = u[F - 1];
#pragma unroll
                    for (int rc = 1; rc < RC; ++rc)
                        if (rc == orc) res = u[(rc + 1) * F - 1];
                    if (p.pairs == PAIRS_SYM) {
                        int a, b;
                        job_decode(p, p.job0 + sjob, a, b);
                        if (p.n_peer > 0) {
                            // both mirror entries of every rank's copy of G (peer memory)
#pragma unroll 1
                            for (int q = 0; q < p.n_peer; ++q) {
                                p.out_peer[q][(long)a * p.B + b] = res;
                                p.out_peer[q][(long)b * p.B + a] = res;
                            }
                        } else {
                            p.out[(long)a * p.B + b] = res;
                            p.out[(long)b * p.B + a] = res;
                        }
                    } else if (p.n_peer > 0) {
                        // one 8-byte store per rank: this pair's entry of every rank's copy of G (peer memory)
#pragma unroll 1
                        for (int q = 0; q < p.n_peer; ++q) p.out_peer[q][p.job0 + sjob] = res;
                    } else {
                        p.out[p.job0 + sjob] = res;   // GRAM: job = a * B + b; BATCH: job = a
                    }
                }
            }
            if (cc == N - 1) {
                if (UFLUSH) {
                    double2* gs = reinterpret_cast<double2*>(gst + (size_t)((w - 1) & p.fbuf_mask) * FBUF) + glane;
#pragma unroll
                    for (int rc = 0; rc < RC; ++rc)
#pragma unroll
                        for (int i = 0; i < DP2; ++i) {
                            gs[(rc * DP2 + i) * GL] = make_double2(ga[GREG ? rc : 0][GREG ? 2 * i : 0], ga[GREG ? rc : 0][GREG ? 2 * i + 1 : 0]);
                            ga[GREG ? rc : 0][GREG ? 2 * i : 0] = 0.0;
                            ga[GREG ? rc : 0][GREG ? 2 * i + 1 : 0] = 0.0;
                        }
                    if (RSYM) {
                        // the production stream's rows are those of the pair the stencil stream starts now
#pragma unroll
                        for (int rc = 0; rc < RC; ++rc)
#pragma unroll
                            for (int i = 0; i < DP2; ++i) xss[RSYM ? rc * DP2 + i : 0][RSYM ? glane : 0] = xr[XREG ? rc : 0][XREG ? i : 0];
                    }
                    syo = yo;
                    syp = reinterpret_cast<const double*>(reinterpret_cast<const char*>(p.Yp) + syo);
                }
                if (REVX && !UFLUSH) {
                    // the pair is complete for this lane: emit its node rows (reversed order), clear the accumulators
                    const long pi = p.job0 + sjob;                      // GRAM: a * B + b; BATCH: a
                    // REV_RECON: fused loss head -- d loss / d X_a += coef * d k(X_a, Y_b) / d X_a (coef and the rows of X_a
                    // were looked up when the stencil stream moved on to this pair)
                    const double coef = cur_coef;
                    double* gx = (RECON && sjob >= 0) ? cur_gx : nullptr;
                    const double* sxb = reinterpret_cast<const double*>(reinterpret_cast<const char*>(p.Xp) + sxo);
                    double* gpair = p.grad ? p.grad + pi * (long)(M * D) : nullptr;     // one 64-bit product per pair, int offsets per row
#pragma unroll
                    for (int rc = 0; rc < RC; ++rc) {
                        const int np = pl * RC + rc;                    // reversed node row
                        double* acc = gacc + (rc * (D + 1)) * GL + glane;
                        if (sjob >= 0 && np < M) {
                            const int go = (M - 1 - np) * D;
                            double* gout = gpair ? gpair + go : nullptr;
                            double* gxr = gx ? gx + go : nullptr;
                            const double* xrw = sxb + np * Dp;
                            if (GREG) {
                                const double sW = ga[GREG ? rc : 0][0];
#pragma unroll
                                for (int k = 0; k < Dp - 1; ++k)
                                    if (k < D) {
                                        const double gy = ga[GREG ? rc : 0][GREG ? k + 1 : 0];
                                        const double gv = KIND == KIND_RBF ? p.inv_kscale * fma(p.gscale, gy, -(__ldg(xrw + 1 + k) * sW)) : p.gscale * gy;
                                        if (!RECON || gout) gout[k] = gv;
                                        if (RECON && gxr && coef != 0.0) atomicAdd(gxr + k, coef * gv);
                                    }
                            } else {
                                const double sW = acc[0];
                                for (int k = 0; k < D; ++k) {
                                    const double gy = acc[(k + 1) * GL];
                                    const double gv = KIND == KIND_RBF ? p.inv_kscale * fma(p.gscale, gy, -(__ldg(xrw + 1 + k) * sW)) : p.gscale * gy;
                                    if (!RECON || gout) gout[k] = gv;
                                    if (RECON && gxr && coef != 0.0) atomicAdd(gxr + k, coef * gv);
                                }
                            }
                        }
                        if (GREG) {
#pragma unroll
                            for (int e2 = 0; e2 < (GREG ? Dp : 1); ++e2) ga[GREG ? rc : 0][e2] = 0.0;
                        } else {
                            for (int k = 0; k <= D; ++k) acc[k * GL] = 0.0;
                        }
                    }
                    sxo = xo; syo = yo;
                    syp = reinterpret_cast<const double*>(reinterpret_cast<const char*>(p.Yp) + syo);
                }
                // node column N-1 has no coarse column: the step computed garbage; re-arm the boundary
                // u[., 0] = 1 and hand the stencil stream the pair the production stream is in
#pragma unroll
                for (int r = 0; r < R; ++r) u[r] = 1.0;
                topprev = 1.0;
                if (RECON) {
                    // first column of the rebuilt solution: u[., NN] of the pair the stencil stream moves on to (staged when
                    // the production column wrapped), and that pair's last row for lane 0
                    cp_async_wait<0>();
                    topprevb = bstg[R * GL + glane];
                    if (!UFLUSH) chk_tol = p.recon_tol * fmax(1.0, fabs(bstg[(R + 1) * GL + glane]));
#pragma unroll
                    for (int r = 0; r < R; ++r) ub[RECON ? r : 0] = bstg[(R - 1 - r) * GL + glane];
                    cur_coef = 0.0;
                    cur_gx = nullptr;
                    if (!UFLUSH && p.gradX != nullptr && pjob >= 0) {
                        const int a = pa;
                        const int b = p.pairs == PAIRS_BATCH ? a : (int)(p.job0 + pjob - (long)a * p.B);
                        cur_coef = p.gout ? __ldg(p.gout + (p.job0 + pjob)) : (a == b ? p.w_diag : p.w_off);
                        cur_gx = p.gradX + (long)a * M * D;
                    }
                }
                c = 0;
                sjob = pjob;
                if (EMIT) ebr = p.brow + (p.job0 + (pjob >= 0 ? pjob : 0)) * p.brow_stride + 1;
            }
            if (cc == pc) {
                // the production column wraps: next pair
                ++w;
                int4 ent = make_int4(-1, (int)xo, (int)yo, pa);
                if (pl == 0) {
                    int job = job_next;
                    if (job < p.njobs) {
                        job_next = (int)(G + atomicAdd(p.counter, 1u));
                        int a, b;
                        job_decode(p, p.job0 + job, a, b);
                        ent.y = (int)((unsigned)a * xstride);
                        ent.z = (int)((unsigned)b * ystride);
                        ent.w = a;
                    } else {
                        job = -1;
                    }
                    ent.x = job;
                    ring_s[sid][w & (RING - 1)] = ent;
                    if (RECON) {
                        // lane 0 looks the pair's boundary arrays up for everybody
                        pa = ent.w;
                        pair_boundaries(job);
                        ring_b[RECON ? sid : 0][RECON ? (w & (RING - 1)) : 0] = make_longlong2((long long)nbr, (long long)nbc);
                    }
                } else if (w >= 0) {
                    ent = ring_s[sid][w & (RING - 1)];     // written >= 1 step (= 1 barrier) ago
                    if (RECON) {
                        const longlong2 eb = ring_b[RECON ? sid : 0][RECON ? (w & (RING - 1)) : 0];
                        nbr = reinterpret_cast<const double*>(eb.x);
                        nbc = reinterpret_cast<const double*>(eb.y);
                    }
                }
                pjob = ent.x;
                if (w >= 0 && ent.x < 0) done = true;
                xo = (unsigned)ent.y;
                yo = (unsigned)ent.z;
                if (RECON) pa = ent.w;
                set_pair();                       // virtual / past the end: the same pair's data again
                if (RECON) stage_first_column(pjob >= 0);
            }
        }
        if (UFLUSH) {
            // ---- 5b. flush: the last lane of the warp has just parked pair fw -- the warp emits its gradient rows ----
            const int fwl = __shfl_sync(FULL, cc == N - 1 ? w : 0, LPP - 1, LPP);   // (w >= 1 once a real pair is complete)
            if (fwl > 0) {
                const int fw = fwl - 1;
                const int4 ent = ring_s[sid][fw & (RING - 1)];
                if (ent.x >= 0) {
                    const long pi = p.job0 + ent.x;                      // GRAM: a * B + b; BATCH: a
                    const int a = ent.w;
                    // (SYM: the pairs a <= b row by row, pi = a A - a (a - 1) / 2 + (b - a))
                    const int b = p.pairs == PAIRS_BATCH ? a
                                  : (RSYM ? a + (int)(pi - ((long)a * p.A - (long)a * (a - 1) / 2)) : (int)(pi - (long)a * p.B));
                    // fused loss head: d loss / d X_a += coef * d k(X_a, Y_b) / d X_a
                    const double coef = p.gradX == nullptr ? 0.0 : (p.gout ? __ldg(p.gout + (RSYM ? (long)a * p.B + b : pi)) : (a == b ? p.w_diag : p.w_off));
                    double* gx = p.gradX ? p.gradX + (long)a * (M * D) : nullptr;
                    // per-pair gradient rows (A, B, M, D): the unordered-pair sweep fills (a, b) here and (b, a) from its column sums
                    double* gpair = p.grad ? p.grad + (RSYM ? (long)a * p.B + b : pi) * (long)(M * D) : nullptr;
                    const double* sxb = reinterpret_cast<const double*>(reinterpret_cast<const char*>(p.Xp) + (unsigned)ent.y);
                    const double* fs = gst + (size_t)(fw & p.fbuf_mask) * FBUF + glane;
                    const double2* fs2 = reinterpret_cast<const double2*>(gst + (size_t)(fw & p.fbuf_mask) * FBUF) + glane;
                    // the rebuilt grid must end at the boundary u[., 0] = 1: a miss beyond recon_tol * max(1, |k|) (errors scale
                    // with the size of the solution; or a NaN) sends the whole call to the stored-grid kernels queued behind
                    // this one.  (A pair that contributes nothing -- zero weight in the loss head, no per-point gradients
                    // asked for -- may miss: the diagonal of Gram(X, X) grows by orders of magnitude more than the rest and
                    // carries no weight in the MMD and the scoring rules.)
                    if (p.grad != nullptr || p.gradX == nullptr || coef != 0.0) {
                        long sl = pi;
                        bool swp = false;
                        if (p.bsym) {
                            swp = a > b;
                            const long lo = swp ? b : a, hi = swp ? a : b;
                            sl = lo * p.A - lo * (lo - 1) / 2 + (hi - lo);
                        }
                        const double kab = swp ? __ldg(p.brow + sl * p.brow_stride + NNf) : __ldg(p.bcol + sl * p.bcol_stride + MMl);
                        const double tol = p.recon_tol * fmax(1.0, fabs(kab));
                        bool bad = false;
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            if ((long)pl * R + r < MMl) bad = bad || !(fabs(fs[(RC * Dp + r) * GL] - 1.0) <= tol);
                        if (bad) *p.flag = 1u;
                    }
#pragma unroll
                    for (int rc = 0; rc < RC; ++rc) {
                        const int np = pl * RC + rc;                    // reversed node row
                        if (np < M && (gpair != nullptr || coef != 0.0)) {
                            const int go = (M - 1 - np) * D;
                            const double* xrw = sxb + np * Dp;
                            double gsum[Dp];                            // [0]: sum of W; [1 + k]: sum of W y_k
#pragma unroll
                            for (int i = 0; i < DP2; ++i) {
                                const double2 v = fs2[(rc * DP2 + i) * GL];
                                gsum[2 * i] = v.x;
                                gsum[2 * i + 1] = v.y;
                            }
                            const double sW = gsum[0];
#pragma unroll
                            for (int k = 0; k < Dp - 1; ++k)
                                if (k < D) {
                                    const double gy = gsum[k + 1];
                                    const double gv = KIND == KIND_RBF ? p.inv_kscale * fma(p.gscale, gy, -(__ldg(xrw + 1 + k) * sW)) : p.gscale * gy;
                                    if (gpair) gpair[go + k] = gv;
                                    if (coef != 0.0) atomicAdd(gx + go + k, coef * gv);
                                }
                        }
                    }
                    if (RSYM) {
                        // the ordered pair (b, a): d loss / d X_b += coef(b, a) * d k(X_b, X_a) / d X_b, and d k(X_b, X_a) / d X_b
                        // = d k(X_a, X_b) / d X_b comes from the column sums the last lane collected (node columns are
                        // reversed like the rows: column q is point N - 1 - q of X_b).  The diagonal pair has no partner.
                        __syncwarp();
                        const double coef2 = (a == b || p.gradX == nullptr) ? 0.0 : (p.gout ? __ldg(p.gout + (long)b * p.B + a) : p.w_off);
                        double* gpair2 = (p.grad && a != b) ? p.grad + ((long)b * p.B + a) * (long)(M * D) : nullptr;
                        if (coef2 != 0.0 || gpair2 != nullptr) {
                            double* gxb = p.gradX + (long)b * (M * D);
                            const double* syb = reinterpret_cast<const double*>(reinterpret_cast<const char*>(p.Yp) + (unsigned)ent.z);
                            for (int q = lane; q < N; q += 32) {
                                const double* part = ycol + (size_t)q * Dp;
                                const double* yrw = syb + q * Dp;
                                const double sW = part[0];
                                for (int k = 0; k < D; ++k) {
                                    const double gv = KIND == KIND_RBF ? p.inv_kscale * fma(-p.gscale * __ldg(yrw + 1 + k), sW, part[1 + k])
                                                                       : p.inv_kscale * part[1 + k];
                                    if (gpair2) gpair2[(N - 1 - q) * D + k] = gv;
                                    if (coef2 != 0.0) atomicAdd(gxb + (N - 1 - q) * D + k, coef2 * gv);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < DP2; ++i) yq[i] = ldg2(yp + 2 * i);
        if (STAGE) {
            // refill the slot just consumed with the rows of the step DEPTH ahead: column c + DEPTH - 1 of this
            // pair, or -- past the dummy column N-1 -- of the pair the production stream is already in
            const int cT = c + DEPTH - 1;
            stage_issue(sq, cT >= N ? pjob : sjob, cT >= N ? cT - N : (cT == N - 1 ? -1 : cT));
            sq = sq + 1 == DEPTH ? 0 : sq + 1;
        }
    };

    int qrun = 0;
#pragma unroll 1
    while (true) {
        const bool alive = !done || sjob >= 0;
        if (NW > 1) {
            if (!__syncthreads_or(alive)) break;
        } else {
            if (!__any_sync(FULL, alive)) break;
        }
        if (UNR == 3) {
            step(std::integral_constant<int, 0>{});
            step(std::integral_constant<int, 1>{});
            step(std::integral_constant<int, 2>{});
        } else {
            step(qrun);
            qrun = qrun == 2 ? 0 : qrun + 1;
        }
    }
    if (MODE == 0 && p.sig_epoch != 0ull) {
        // sharded Gram: the kernel ends with the barrier across the ranks -- the last block to finish (all results of this
        // rank are then on their way to every copy of G) signals every rank and waits for every rank's signal
        if (NW > 1) __syncthreads(); else __syncwarp();
        if (threadIdx.x < 32) {
            unsigned prev = 0u;
            if (threadIdx.x == 0) {
                __threadfence_system();
                prev = atomicAdd(p.counter + 32, 1u);
            }
            prev = __shfl_sync(FULL, prev, 0);
            if (prev == gridDim.x - 1) rank_barrier(p);
        }
    }
}

}  // namespace skb
