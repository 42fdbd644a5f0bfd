// skb_tile.cuh -- the "tile" forward kernel of the fused static kinds (Linear / RBF): the headline path of
// compute_Gram for large batches (BASELINE configs[2] and configs[4]).
//
// Decomposition (round 2; DESIGN.md 3b has the measurements that led here):
//   * a LANE owns one path pair; the 32 lanes of a warp hold 32 pairs of one TILE (32 consecutive rows a of X
//     against one column b of Y, or 32 consecutive batch entries), all at the same grid position;
//   * a WARP owns one STRIP of R = RC * 2^d fine rows of those 32 grids and sweeps it left to right in macro
//     steps of one coarse column (F = 2^d fine columns, R x F cells per lane and step, all in registers);
//   * the W warps of a block own W consecutive strips (a BAND of W R rows) and run as a pipeline: warp w+1 works
//     one or more steps behind warp w and receives the bottom row of warp w's strip (F values per lane and
//     step) plus the static-kernel column difference of the node row the two strips share through a ring in
//     shared memory, guarded by per-warp progress counters (st.release / ld.acquire, no block barrier);
//   * grids taller than one band continue in another JOB = (tile, band): the bottom row of the band goes to global
//     memory (coalesced: [column][value][lane]) and is read back by the job (tile, band + 1) -- on whichever block
//     pops it from the queue.  Jobs are queued band-major, so the row a job depends on is complete long before the
//     job is popped; any path length is covered this way;
//   * a HELPER warp (warp W of the block, no stencil work) keeps the W stencil warps identical: it pops the job
//     queue, feeds the inbound ring of stencil warp 0 from global memory with cp.async several steps ahead (band
//     0: u = 1 and the column differences of node row 0; band > 0: the row the band above left), and drains the
//     outbound ring of the last stencil warp to global memory.  (Measured: with warp 0 doing its own feeding it
//     ran 12 % more instructions than the others and, being the head of the pipeline, paced all of them.)
// What this buys over fwd5_kernel (one pair per half-warp, lanes skewed by one step): every lane of a warp is
// at the same column of the same job, so the per-pair events (output, boundary re-arm, next pair's x rows) are
// warp-uniform branches taken once per 64 steps instead of divergent blocks run in half of the steps; there is
// no virtual pair, no dummy stencil step (the step without a coarse column runs the production only), no warp
// barrier; the y row of a step is one broadcast load.  Per step and lane: 264 DP instructions out of ~310.
//
// Static kernel: as in fwd5 (column differences d[i][j] = k[i][j+1] - k[i][j] pre-scaled by 4^-d / sqrt(12); the
// increment of a coarse cell is one subtraction), but the exp argument arrives pre-scaled by 2048 / ln 2 (folded
// into the prepared rows), so the range reduction is three additions (no hi/lo split of ln 2), and a 2^11-entry
// table leaves a degree-3 polynomial: 7 DP instructions per exp instead of 9.
//
// Reference semantics replaced: sigkernel/cuda_backend.py:121-160 (+ :6-49), static_kernels.py:17-33, 42-73,
// sigkernel.py:362-364, 607-613.  fp64 throughout, FMA arithmetic, results within 1e-12 of the oracle.
#pragma once
#include "skb_solver.cuh"

namespace skb {

constexpr int TTAB = 2048;                 // exp table: kscale * 2^(j / 2048)
// depth of the inter-warp rings (steps); warp 0 prefetches its boundary RD - 1 steps ahead.  The tile path needs
// N - 1 >= TILE_RD_MAX + 1 coarse columns per job (tile_applies() asks for len_y >= 16).
constexpr int TILE_RD_MAX = 8;
__host__ __device__ constexpr int tile_ring_depth(int rc, int dp2) { return rc * dp2 <= 20 ? 8 : 2; }
constexpr int TILE_JR = 16;                // depth of the job ring

// Arguments of tile_fwd_kernel.  Plain data, passed by value.
struct TArgs {
    const double* Xp;      // prepared X rows [A*M][Dp]: (s*nx, s*c*x_0 ..., 0 pad), s = 2048 / ln 2 (RBF) or kscale (Linear)
    const double* Yp;      // prepared Y rows [B*N][Dp]: (s*ny, y_0 ..., 0 pad)
    double* out;           // k(X_a, Y_b): GRAM (A,B) row-major, BATCH (A,)
    const double* d0;      // [ntiles][N-1][32]: column differences of node row 0 (top boundary of band 0)
    double2* bnd;          // [ntiles][nbands-1][N-1][F/2+1][32]: bottom row (pairs of fine columns) + (d of the last node
                           // row, -) of every band but the last: the layout of a ring slot, copied with cp.async
    unsigned int* ready;   // [ntiles][nbands-1]: columns of bnd published so far
    unsigned int* counter; // job queue
    int A, B, M, N;
    int pairs;             // PAIRS_GRAM or PAIRS_BATCH
    int nta;               // GRAM: tiles per column of Y = ceil(A / 32)
    int ntiles, nbands, njobs;   // job = band * ntiles + tile (band-major)
    double kscale;         // 4^-d / sqrt(12)
    double sqrt3;
    double c1, c2, c3;     // e^(r c) - 1 = r (c1 + r (c2 + r c3)), c = ln 2 / 2048
};

__device__ __forceinline__ unsigned ld_acquire_cta(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta(unsigned* p, unsigned v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// exp(xs * ln2 / 2048) * kscale for xs <= ~0 (the table holds kscale * 2^(j/2048)).  The underflow guard clamps
// xs to >= -2031616 (x >= -687.6) through an unsigned min on the high word; NaN passes through.
__device__ __forceinline__ double exp_scaled(double xs, const double* __restrict__ tab, const TArgs& p) {
    const double MAGIC = 6755399441055744.0;             // 1.5 * 2^52
    const unsigned hi = min((unsigned)__double2hiint(xs), 0xC13F0000u);
    const double xc = __hiloint2double((int)hi, __double2loint(xs));
    const double t = xc + MAGIC;
    const double nf = t - MAGIC;                          // nearest integer = 2048 n + j
    const double r = xc - nf;                             // exact, |r| <= 1/2
    double q = fma(r, p.c3, p.c2);
    q = fma(q, r, p.c1);
    q = q * r;                                            // e^(r ln2/2048) - 1, truncation 3.4e-17
    const int ti = __double2loint(t);
    const double tj = tab[ti & (TTAB - 1)];
    const double v = fma(tj, q, tj);
    return __hiloint2double(__double2hiint(v) + (ti & ~(TTAB - 1)) * 512, __double2loint(v));
}

// 16-byte / 8-byte asynchronous copies global -> shared (LDGSTS); the 16-byte form bypasses L1
__device__ __forceinline__ void tile_cp16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void tile_cp8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void tile_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void tile_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

template <int KIND, int RC, int LOGD, int DP2, int W>
__global__ void __launch_bounds__(32 * (W + 1), 1) tile_fwd_kernel(const TArgs p) {
    constexpr int F = 1 << LOGD, R = RC * F, Dp = 2 * DP2, H = (F + 1) / 2, HS = H + 1;
    constexpr int RD = tile_ring_depth(RC, DP2), JR = TILE_JR;
    constexpr bool XREG = (RC * DP2 <= 20);       // x rows of the lane's pair in registers, else in shared memory
    constexpr int NX = RC * DP2;                   // 16-byte pieces of x rows per lane
    constexpr int NT = 32 * W;                     // stencil threads
    static_assert(F >= 2, "ring slots hold pairs of fine columns");
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int N = p.N, M = p.M;
    const int NS = N - 1;                          // stencil steps (coarse columns) per job
    const int MMf = (M - 1) << LOGD;               // fine rows of the grid

    // shared memory: exp table | ring[W+1][RD][HS][32] double2 | x rows: staging of the next job (XREG) or two
    // buffers (flip per job) | prog[W+2] | jobring[JR].  ring[w] is the inbound ring of stencil warp w (ring[0] is fed
    // by the helper), ring[W] the outbound ring of the last one.  A slot = H pairs of bottom-row values + (d, -).
    // prog[0]: columns fed by the helper; prog[w+1]: steps finished by stencil warp w; prog[W+1]: steps drained.
    extern __shared__ double smem_tile[];
    double* const etab = smem_tile;
    double2* const ring = reinterpret_cast<double2*>(smem_tile + (KIND == KIND_RBF ? TTAB : 0));
    double2* const xs_s = ring + (W + 1) * RD * HS * 32;
    unsigned* const prog = reinterpret_cast<unsigned*>(xs_s + (XREG ? 1 : 2) * NX * NT);
    int* const jobring = reinterpret_cast<int*>(prog + W + 2);

    if (KIND == KIND_RBF) {
        for (int j = tid; j < TTAB; j += 32 * (W + 1)) etab[j] = p.kscale * exp2((double)j * (1.0 / TTAB));
    }
    if (tid < W + 2) prog[tid] = 0u;
    if (tid == 0) jobring[0] = (int)blockIdx.x < p.njobs ? (int)blockIdx.x : -1;
    __syncthreads();

    int sjob = jobring[0];
    if (sjob < 0) return;

    if (w == W) {
        // ======================= helper warp: job queue, feed of ring[0], drain of ring[W] =======================
        auto pop = [&]() {
            int nj = 0;
            if (lane == 0) nj = (int)(gridDim.x + atomicAdd(p.counter, 1u));
            nj = __shfl_sync(FULL, nj, 0);
            return nj < p.njobs ? nj : -1;
        };
        int fj = 0, fjob = sjob, fc = 0;        // feed: job index / id / next column
        unsigned gf = 0u;                        // feed: global index of the next column (the numbering of the stencil warps)
        int dj = 0, djob = sjob, dc = 0;        // drain: job index / id / next column
        unsigned gd = 0u;
        unsigned seen0 = 0u, seenL = 0u, avail = 0u;
        {
            const int nj = pop();
            if (lane == 0) jobring[1] = nj;      // released with the first fed column
        }
        while (fjob >= 0 || djob >= 0) {
            bool progress = false;
            if (fjob >= 0) {
                const int band = fjob / p.ntiles, tile = fjob - band * p.ntiles;
                int can = NS - fc;
                if (can > 4) can = 4;
                if (gf + (unsigned)can > seen0 + (unsigned)RD) {
                    seen0 = ld_acquire_cta(&prog[1]);
                    const int room = (int)(seen0 + (unsigned)RD - gf);
                    if (can > room) can = room;
                }
                const unsigned* rdy = nullptr;
                if (band > 0 && can > 0) {
                    rdy = p.ready + (size_t)tile * (p.nbands - 1) + (band - 1);
                    if (avail < (unsigned)(fc + can)) {
                        avail = ld_acquire_gpu(rdy);
                        if ((int)avail - fc < can) can = (int)avail - fc;
                    }
                }
                if (can > 0) {
                    for (int k = 0; k < can; ++k) {
                        double2* dst = ring + ((0 * RD + (int)((gf + (unsigned)k) & (RD - 1))) * HS) * 32 + lane;
                        if (band == 0) {
#pragma unroll
                            for (int h = 0; h < H; ++h) dst[h * 32] = make_double2(1.0, 1.0);
                            tile_cp8(dst + H * 32, p.d0 + ((size_t)tile * NS + (fc + k)) * 32 + lane);
                        } else {
                            const double2* src = p.bnd + ((((size_t)tile * (p.nbands - 1) + (band - 1)) * NS) + (fc + k)) * (HS * 32) + lane;
#pragma unroll
                            for (int h = 0; h < HS; ++h) tile_cp16(dst + h * 32, src + h * 32);
                        }
                    }
                    tile_cp_commit();
                    tile_cp_wait<0>();
                    __threadfence_block();
                    __syncwarp();
                    gf += (unsigned)can;
                    fc += can;
                    if (lane == 0) st_release_cta(&prog[0], gf);
                    progress = true;
                    if (fc == NS) {
                        // next job: its id was popped one job ago; pop the one after it now
                        ++fj;
                        fjob = jobring[fj & (JR - 1)];
                        fc = 0;
                        avail = 0u;
                        if (fjob >= 0) {
                            const int nj = pop();
                            if (lane == 0) jobring[(fj + 1) & (JR - 1)] = nj;     // released with the next fed column
                        }
                    }
                }
            }
            if (djob >= 0) {
                const int band = djob / p.ntiles, tile = djob - band * p.ntiles;
                if (band + 1 >= p.nbands) {
                    // nothing leaves the last band: skip the job -- once the entry of the job after it exists (entry k
                    // is written when the feed side reaches job k - 1; the drain never passes the feed side)
                    if (fj > dj || fjob < 0) {
                        gd += (unsigned)NS;
                        ++dj;
                        djob = jobring[dj & (JR - 1)];
                        if (lane == 0) st_release_cta(&prog[W + 1], gd);
                        progress = true;
                    }
                } else {
                    if (seenL < gd + 1u) seenL = ld_acquire_cta(&prog[W]);
                    if (seenL >= gd + 1u) {
                        const double2* src = ring + ((W * RD + (int)(gd & (RD - 1))) * HS) * 32 + lane;
                        double2* dst = p.bnd + ((((size_t)tile * (p.nbands - 1) + band) * NS) + dc) * (HS * 32) + lane;
#pragma unroll
                        for (int h = 0; h < HS; ++h) __stcg(dst + h * 32, src[h * 32]);
                        ++gd;
                        ++dc;
                        __syncwarp();
                        if (lane == 0) st_release_cta(&prog[W + 1], gd);
                        if ((dc & 7) == 0 || dc == NS) {
                            asm volatile("fence.acq_rel.gpu;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) st_release_gpu(p.ready + (size_t)tile * (p.nbands - 1) + band, (unsigned)dc);
                        }
                        progress = true;
                        if (dc == NS) {
                            dc = 0;
                            ++dj;
                            // (entry dj was written when the feed side reached job dj - 1, which the stencil warps have finished)
                            djob = jobring[dj & (JR - 1)];
                        }
                    }
                }
            }
            if (!progress) __nanosleep(64);
        }
        return;
    }

    // ======================================== stencil warps (all identical) ========================================
    // ---- production stream ------------------------------------------------------------------------------
    double2 xr[XREG ? RC : 1][DP2];
    int xbuf = 0;                                  // !XREG: buffer the production reads
    const double* ybase = p.Yp;
    // issue the loads of a job's x rows (this warp's strip) into the staging buffer / the other buffer
    auto stage_x = [&](int job, int buf) {
        const int band = job / p.ntiles, tile = job - band * p.ntiles;
        int a;
        if (p.pairs == PAIRS_BATCH) a = tile * 32 + lane;
        else a = (tile - (tile / p.nta) * p.nta) * 32 + lane;
        a = a < p.A ? a : p.A - 1;
        const int s = band * W + w;
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) {
            int row = s * RC + 1 + rc;             // the strip's own node rows (the one above comes from warp w-1)
            row = row < M ? row : M - 1;
            const double* xp = p.Xp + ((size_t)a * M + row) * Dp;
#pragma unroll
            for (int i = 0; i < DP2; ++i) tile_cp16(xs_s + ((buf * NX + rc * DP2 + i) * NT + tid), xp + 2 * i);
        }
        tile_cp_commit();
    };
    // make the staged rows current (all async copies of this thread have landed)
    auto take_x = [&](int job, int buf) {
        const int tile = job % p.ntiles;
        int b;
        if (p.pairs == PAIRS_BATCH) b = tile * 32 + lane;
        else b = tile / p.nta;
        b = b < p.B ? b : p.B - 1;
        ybase = p.Yp + (size_t)b * N * Dp;
        tile_cp_wait<0>();
        if (XREG) {
#pragma unroll
            for (int rc = 0; rc < RC; ++rc)
#pragma unroll
                for (int i = 0; i < DP2; ++i) xr[XREG ? rc : 0][i] = xs_s[(rc * DP2 + i) * NT + tid];
        } else {
            xbuf = buf;
        }
    };
    double2 yq[DP2];
    auto load_y = [&](int col) {
        const double* yp = ybase + (size_t)col * Dp;
#pragma unroll
        for (int i = 0; i < DP2; ++i) yq[i] = ldg2(yp + 2 * i);
    };
    double klast[RC], dcur[RC];
    auto produce = [&](double* knew) {
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) {
            double2 xv = XREG ? xr[XREG ? rc : 0][0] : xs_s[(xbuf * NX + rc * DP2) * NT + tid];
            double acc = fma(xv.y, yq[0].y, xv.x + yq[0].x);
#pragma unroll
            for (int i = 1; i < DP2; ++i) {
                xv = XREG ? xr[XREG ? rc : 0][i] : xs_s[(xbuf * NX + rc * DP2 + i) * NT + tid];
                acc = fma(xv.y, yq[i].y, fma(xv.x, yq[i].x, acc));
            }
            knew[rc] = KIND == KIND_RBF ? exp_scaled(acc, etab, p) : acc;
        }
    };

    int pjob = sjob;          // job of the production stream (runs two node columns ahead of the stencil)
    stage_x(pjob, 0);
    take_x(pjob, 0);
    {
        // prologue: node columns 0 and 1 of the first job -> d[0]
        load_y(0);
        produce(klast);
        load_y(1);
        double k1[RC];
        produce(k1);
#pragma unroll
        for (int rc = 0; rc < RC; ++rc) { dcur[rc] = k1[rc] - klast[rc]; klast[rc] = k1[rc]; }
        load_y(2);
    }
    int e = 2;                // next production column (N >= 4)
    int njob = -1;            // the job after sjob (known from c == 1 on)

    unsigned seen_up = 0u, seen_dn = 0u;
    unsigned g = 0u;          // this warp's step counter over the whole kernel (the same numbering in every warp)
    int jn = 0;               // jobs whose stencil this warp has finished
    const unsigned* const pup = &prog[w];          // upstream: helper (w == 0) or stencil warp w-1
    unsigned* const pown = &prog[w + 1];
    const unsigned* const pdn = &prog[w + 2];      // downstream: stencil warp w+1 or the helper's drain counter
    const double2* const rin = ring + (w * RD * HS) * 32 + lane;
    double2* const rout = ring + ((w + 1) * RD * HS) * 32 + lane;

    double u[R];
    while (sjob >= 0) {
        const int band = sjob / p.ntiles, tile = sjob - band * p.ntiles;
        const int s = band * W + w;
        const bool hand_down = (w < W - 1) || (band + 1 < p.nbands);
#pragma unroll
        for (int r = 0; r < R; ++r) u[r] = 1.0;
        double topprev = 1.0;
        njob = -1;

#pragma unroll 1
        for (int c = 0; c < NS; ++c, ++g) {
            // ---- 1. the row above the strip ---------------------------------------------------------------
            const int slot = (int)(g & (RD - 1));
            if (seen_up < g + 1u) {
                do { seen_up = ld_acquire_cta(pup); } while (seen_up < g + 1u);
            }
            if (c == 1) {
                // the id of the next job of this block was published before the first column of this job was
                // fed / handed down; start the loads of its x rows
                njob = jobring[(jn + 1) & (JR - 1)];
                if (njob >= 0) stage_x(njob, XREG ? 0 : (xbuf ^ 1));
            }
            double tops[F], dtop;
            {
                const double2* ru = rin + (slot * HS) * 32;
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const double2 v = ru[h * 32];
                    tops[2 * h] = v.x;
                    tops[2 * h + 1] = v.y;
                }
                dtop = ru[H * 32].x;
            }
            // ---- 2. stencil coefficients of coarse column c -----------------------------------------------
            // e = g / sqrt(12) (g = the refined increment):  -b = e^2 - 1,  a = 1 + g/2 + g^2/12 = sqrt(3) e + (2 - b)
            double ca[RC], cb[RC];
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) {
                const double el = dcur[rc] - (rc == 0 ? dtop : dcur[rc > 0 ? rc - 1 : 0]);
                cb[rc] = fma(el, el, -1.0);
                ca[rc] = fma(el, p.sqrt3, cb[rc] + 2.0);
            }
            // ---- 3. the stencil: R rows x F fine columns in registers, anti-diagonal order -----------------
            double U[R][F];
#pragma unroll
            for (int dgl = 0; dgl < R + F - 1; ++dgl) {
                double ss[F], tt[F];
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    const int r = dgl - f;
                    if (r >= 0 && r < R) {
                        const int rm = r > 0 ? r - 1 : 0, fm = f > 0 ? f - 1 : 0;
                        const double left = f == 0 ? u[r] : U[r][fm];
                        const double up = r == 0 ? tops[f] : U[rm][f];
                        const double diag = r == 0 ? (f == 0 ? topprev : tops[fm]) : (f == 0 ? u[rm] : U[rm][fm]);
                        ss[f] = left + up;
                        tt[f] = cb[r >> LOGD] * diag;
                    }
                }
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    const int r = dgl - f;
                    if (r >= 0 && r < R) U[r][f] = fma(ca[r >> LOGD], ss[f], tt[f]);
                }
            }
            topprev = tops[F - 1];
#pragma unroll
            for (int r = 0; r < R; ++r) u[r] = U[r][F - 1];
            // ---- 4. hand the bottom row (and d of the strip's last node row) to the strip below -------------
            if (hand_down) {
                if (g >= (unsigned)RD && seen_dn < g - (unsigned)RD + 1u) {
                    do { seen_dn = ld_acquire_cta(pdn); } while (seen_dn < g - (unsigned)RD + 1u);
                }
                double2* wu = rout + (slot * HS) * 32;
#pragma unroll
                for (int h = 0; h < H; ++h) wu[h * 32] = make_double2(U[R - 1][2 * h], U[R - 1][2 * h + 1]);
                wu[H * 32] = make_double2(dcur[RC - 1], 0.0);
            }
            __syncwarp();
            if (lane == 0) st_release_cta(pown, g + 1u);
            // ---- 5. production: node column e of the production job -> d of the next step -------------------
            {
                double knew[RC];
                produce(knew);
#pragma unroll
                for (int rc = 0; rc < RC; ++rc) { dcur[rc] = knew[rc] - klast[rc]; klast[rc] = knew[rc]; }
            }
            if (++e == N) {
                // the production stream moves on to the next job of this block (its x rows were staged at c == 1)
                e = 0;
                pjob = njob;
                if (njob >= 0) take_x(njob, XREG ? 0 : (xbuf ^ 1));   // past the end: the same pair's data again (unused)
            }
            load_y(e);
        }
        // ---- 6. output: u[MM, NN] sits in the strip that holds fine row MM - 1 ------------------------------
        {
            const int rstar = (MMf - 1) - s * R;
            if (rstar >= 0 && rstar < R) {
                double res = u[0];
#pragma unroll
                for (int r = 1; r < R; ++r)
                    if (r == rstar) res = u[r];
                if (p.pairs == PAIRS_BATCH) {
                    const int a = tile * 32 + lane;
                    if (a < p.A) p.out[a] = res;
                } else {
                    const int b = tile / p.nta;
                    const int a = (tile - b * p.nta) * 32 + lane;
                    if (a < p.A) p.out[(size_t)a * p.B + b] = res;
                }
            }
        }
        // ---- 7. the step without a coarse column: production only (node column 1 of the next job -> its d[0]) --
        {
            double knew[RC];
            produce(knew);
#pragma unroll
            for (int rc = 0; rc < RC; ++rc) { dcur[rc] = knew[rc] - klast[rc]; klast[rc] = knew[rc]; }
            ++e;
            load_y(e);
        }
        ++jn;
        sjob = pjob;
    }
    tile_cp_wait<0>();
}

}  // namespace skb
