// skb_api.cu -- the extern "C" boundary declared in include/sigkernel_b200.h.
// Argument validation, workspace carving and kernel dispatch only; no allocation, no sync.
#include <string.h>
#include "skb_common.cuh"
#include "skb_host.h"

namespace skb {

constexpr int MODE_FWD = 0, MODE_FWD_STORE = 1, MODE_REV_S = 2, MODE_REV_GRAD = 3, MODE_REV_RECON = 4, MODE_FWD_EMIT = 5,
              MODE_REV_RECON_SYM = 6;

static thread_local int g_last_cuda = 0;

int check_cuda(cudaError_t e) {
    if (e == cudaSuccess) return SKB_OK;
    g_last_cuda = (int)e;
    return SKB_ERR_CUDA;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static const size_t kCounterBytes = 256;
static const size_t kScratchBudget = (size_t)8 << 30;

static int check_common(int A, int B, int M, int N, int dyadic_order, int scheme, int pairs, int arith) {
    if (A <= 0 || B <= 0 || M < 2 || N < 2 || dyadic_order < 0 || dyadic_order > 20) return SKB_ERR_BAD_SHAPE;
    if (scheme != SKB_SCHEME_S2 && scheme != SKB_SCHEME_S1) return SKB_ERR_BAD_ENUM;
    if (pairs != SKB_PAIRS_GRAM && pairs != SKB_PAIRS_BATCH && pairs != SKB_PAIRS_SYM) return SKB_ERR_BAD_ENUM;
    if (arith != SKB_ARITH_FMA && arith != SKB_ARITH_EXACT) return SKB_ERR_BAD_ENUM;
    if (pairs != SKB_PAIRS_GRAM && A != B) return SKB_ERR_BAD_SHAPE;
    if (pairs == SKB_PAIRS_SYM && M != N) return SKB_ERR_BAD_SHAPE;
    return SKB_OK;
}

static long njobs_of(int A, int B, int pairs) {
    if (pairs == SKB_PAIRS_GRAM) return (long)A * B;
    if (pairs == SKB_PAIRS_BATCH) return A;
    return (long)A * (A + 1) / 2;
}

// prepped-path scale factors: k = exp(nx + ny + <c x, y>) (RBF) or <c x, y> (Linear)
static void prep_factors(int kind, double param, double& cx, double& nsc) {
    if (kind == SKB_STATIC_RBF) {
        cx = 2.0 / param;
        nsc = -1.0 / param;
    } else {
        cx = param;
        nsc = 0.0;
    }
}

static double scale4_of(int d) { return 1.0 / (double)(1ull << (2 * d)); }

// bytes of forward grid one pair needs in the adjoint scratch, and the front pad
static size_t grid_doubles_per_pair(int M, int N, int d) {
    const int R = solver_rows_per_lane(M, d);
    if (R < 0) return 0;
    return (size_t)((long)(N - 1) << d) * (size_t)(32L * R);
}
static size_t front_pad_doubles(int M, int d) {
    const int R = solver_rows_per_lane(M, d);
    return R < 0 ? 0 : (size_t)(32L * R);
}

static KArgs base_args(int A, int B, int M, int N, int d, int scheme, int pairs) {
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.A = A; a.B = B; a.M = M; a.N = N; a.Mv = M; a.Nv = N;
    a.pairs = pairs;
    a.s1 = scheme == SKB_SCHEME_S1;
    a.scale4 = scale4_of(d);
    return a;
}

// ---- generic fallback (any len_x, len_y, dyadic order): coarse increments + row-band sweep ---------
static const size_t kGenericBudget = (size_t)1 << 30;

static size_t generic_doubles_per_job(int M, int N, int d, bool need_static) {
    const size_t NNf = (size_t)(N - 1) << d;
    return (need_static ? (size_t)M * N : 0) + (size_t)(M - 1) * (N - 1) + 2 * NNf;
}

static size_t generic_workspace_bytes(long njobs, int M, int N, int d, bool need_static) {
    const size_t per = generic_doubles_per_job(M, N, d, need_static) * sizeof(double);
    size_t jobs = (size_t)njobs;
    size_t cap = kGenericBudget / per;
    if (cap < 1) cap = 1;
    if (jobs > cap) jobs = cap;
    return align256(jobs * per) + 1024;
}

// src_kind: KIND_RBF / KIND_LINEAR (a.Xp / a.Yp prepared), KIND_STATIC (a.Ks = coarse static matrix) or
// KIND_INC (a.Ks = fine increments, d must be 0).  a.M, a.N are the path lengths (KIND_INC: MM+1, NN+1).
static int run_generic_forward(int src_kind, KArgs a, int d, bool exact, long njobs, char* scratch, size_t scratch_bytes,
                               cudaStream_t st) {
    const int M = a.M, N = a.N;
    const long MMf = (long)(M - 1) << d, NNf = (long)(N - 1) << d;
    if (MMf > 0x3fffffffL || NNf > 0x3fffffffL) return SKB_ERR_BAD_SHAPE;
    const bool need_static = (src_kind == KIND_RBF || src_kind == KIND_LINEAR);
    const bool need_inc = src_kind != KIND_INC;
    const size_t per = generic_doubles_per_job(M, N, d, need_static) * sizeof(double);
    if (scratch_bytes < per + 1024) return SKB_ERR_WORKSPACE;
    long chunk = (long)((scratch_bytes - 1024) / per);
    if (chunk > njobs) chunk = njobs;
    int rcb = 1;
    while (rcb < 8 && 32L * rcb < MMf + 1) rcb <<= 1;
    const long H = 32L * rcb - 1;                         // fine rows per band
    const long nbands = (MMf + H - 1) / H;
    double* base = (double*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    for (long j0 = 0; j0 < njobs; j0 += chunk) {
        const long nj = njobs - j0 < chunk ? njobs - j0 : chunk;
        double* Ks = base;
        double* incc = Ks + (need_static ? (size_t)nj * M * N : 0);
        double* rowA = incc + (size_t)nj * (M - 1) * (N - 1);
        double* rowB = rowA + (size_t)nj * NNf;
        int rc;
        const double* inc_src;
        if (need_static) {
            rc = launch_static_matrix(a, src_kind, j0, nj, Ks, st);
            if (rc) return rc;
            rc = launch_coarse_increments(Ks, incc, nj, M, N, a.scale4, exact, st);
            if (rc) return rc;
            inc_src = incc;
        } else if (need_inc) {
            rc = launch_coarse_increments(a.Ks + (size_t)j0 * M * N, incc, nj, M, N, a.scale4, exact, st);
            if (rc) return rc;
            inc_src = incc;
        } else {
            inc_src = a.Ks + (size_t)j0 * (M - 1) * (N - 1);
        }
        for (long band = 0; band < nbands; ++band) {
            KArgs b = a;
            const long row0 = band * H;
            const long hb = MMf - row0 < H ? MMf - row0 : H;
            b.Ks = inc_src;
            b.counter_clean = 0;                     // one launch per band: each needs its own reset
            b.Xp = b.Yp = nullptr;
            b.M = (int)hb + 1;
            b.N = (int)NNf + 1;
            b.Mv = (int)MMf; b.Nv = (int)NNf;
            b.Mc = M - 1; b.Nc = N - 1;
            b.dshift = d;
            b.band_row0 = (int)row0;
            b.band_top = band > 0 ? ((band & 1) ? rowA : rowB) : nullptr;
            b.band_bot = band + 1 < nbands ? ((band & 1) ? rowB : rowA) : nullptr;
            b.out = band + 1 < nbands ? nullptr : a.out;
            b.job0 = j0;
            b.njobs = (int)nj;
            b.scale4 = 1.0;
            rc = launch_solver(MODE_FWD, KIND_INCV, 0, exact, b, st);
            if (rc) return rc;
        }
    }
    return SKB_OK;
}

// ---- adjoint by reconstruction: layout of the boundary context (last row / last column of every pair's grid) ----
static size_t ctx_row_doubles(int N, int d) { return ((((size_t)(N - 1)) << d) + 1 + 3) & ~(size_t)3; }
static size_t ctx_col_doubles(int M, int d) { return ((((size_t)(M - 1)) << d) + 1 + 3) & ~(size_t)3; }
static const size_t kFlagOffset = 64;            // the reconstruction flag lives in the counter block of the workspace
static const double kReconTol = 1e-10;           // |rebuilt u[., 0] - 1| beyond this sends the call to the stored-grid kernels
static const size_t kFallbackBudget = (size_t)1 << 30;

static void set_ctx(KArgs& a, void* ctx, long njobs, int M, int N, int d) {
    a.brow = (double*)ctx;
    a.brow_stride = (long)ctx_row_doubles(N, d);
    a.bcol = a.brow + (size_t)njobs * a.brow_stride;
    a.bcol_stride = (long)ctx_col_doubles(M, d);
}

// acc += sum_{a,b} w(a,b) G[a,b] with w = w_diag on the diagonal, w_off elsewhere (row-major (A, B)): the reductions behind
// the MMD, the distance and the scoring rules (sigkernel.py:130-197)
__global__ void gram_reduce_kernel(const double* __restrict__ G, long n, int B, double w_diag, double w_off, double* __restrict__ acc) {
    // B == 0: a vector of n batch entries, every one weighted w_diag
    double s = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const bool dg = B == 0 || (i / B == i % B);
        s = fma(dg ? w_diag : w_off, G[i], s);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(acc, s);
}

// gradX = (accumulate ? gradX : 0) + scale * gtmp, scale = out_scale * (out_scale_dev ? *out_scale_dev : 1)
__global__ void combine_kernel(double* __restrict__ gradX, const double* __restrict__ gtmp, size_t n, double out_scale,
                               const double* __restrict__ out_scale_dev, int accumulate) {
    const double sc = out_scale * (out_scale_dev ? *out_scale_dev : 1.0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        gradX[i] = accumulate ? fma(sc, gtmp[i], gradX[i]) : sc * gtmp[i];
}

static int launch_combine(double* gradX, const double* gtmp, size_t n, double out_scale, const double* out_scale_dev, int accumulate,
                          cudaStream_t st) {
    size_t blocks = (n + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    combine_kernel<<<(unsigned)blocks, 256, 0, st>>>(gradX, gtmp, n, out_scale, out_scale_dev, accumulate);
    return check_launch();
}

}  // namespace skb

using namespace skb;

extern "C" {

const char* skb_error_string(int code) {
    switch (code) {
        case SKB_OK: return "ok";
        case SKB_ERR_BAD_SHAPE: return "bad shape (sizes must be positive, M,N >= 2, BATCH/SYM need A == B, SYM needs M == N)";
        case SKB_ERR_BAD_ENUM: return "unknown or unsupported enum value (static kind / scheme / pairs / arith / dtype)";
        case SKB_ERR_WORKSPACE: return "workspace missing or too small (see skb_*_workspace_bytes)";
        case SKB_ERR_UNSUPPORTED: return "shape not instantiated in this build (needs ceil(M/32 rounded to a power of two) * 2^dyadic_order <= 32)";
        case SKB_ERR_CUDA: return "CUDA error (see skb_last_cuda_error)";
        case SKB_ERR_NULL: return "required pointer is NULL";
        default: return "unknown error code";
    }
}

int skb_last_cuda_error(void) { return g_last_cuda; }
int skb_version(void) { return 2; }
void skb_set_warps_per_sm(int warps) { set_warps_per_sm(warps); }
void skb_set_tile_mode(int mode) { set_tile_mode(mode); }
void skb_set_adjoint_mode(int mode) { set_adjoint_mode(mode); }
void skb_set_deriv_mode(int mode) { set_deriv_mode(mode); }
void skb_set_profile_events(void* start_event, void* stop_event) { set_profile_events(start_event, stop_event); }

int skb_forward_plan(int M, int N, int D, int dyadic_order, int static_kind, int scheme) {
    if (M < 2 || N < 2 || D <= 0 || dyadic_order < 0 || dyadic_order > 20) return SKB_ERR_BAD_SHAPE;
    if (static_kind != SKB_STATIC_LINEAR && static_kind != SKB_STATIC_RBF) return SKB_ERR_BAD_ENUM;
    if (scheme != SKB_SCHEME_S2 && scheme != SKB_SCHEME_S1) return SKB_ERR_BAD_ENUM;
    const int kind = static_kind == SKB_STATIC_RBF ? KIND_RBF : KIND_LINEAR;
    if (fwd5_applies(kind, M, N, D, dyadic_order, scheme == SKB_SCHEME_S1)) {
        int lpp = 32;
        const int nw = fwd5_warps_per_pair(M, dyadic_order, &lpp);
        return lpp == 16 ? 4 : (nw == 1 ? 5 : (nw == 2 ? 6 : 7));
    }
    return solver_rows_per_lane(M, dyadic_order) >= 0 ? 1 : 0;
}

int skb_adjoint_plan(int M, int N, int D, int dyadic_order, int static_kind, int scheme) {
    if (M < 2 || N < 2 || D <= 0 || dyadic_order < 0 || dyadic_order > 20) return SKB_ERR_BAD_SHAPE;
    if (static_kind != SKB_STATIC_LINEAR && static_kind != SKB_STATIC_RBF) return SKB_ERR_BAD_ENUM;
    if (scheme != SKB_SCHEME_S2 && scheme != SKB_SCHEME_S1) return SKB_ERR_BAD_ENUM;
    const int kind = static_kind == SKB_STATIC_RBF ? KIND_RBF : KIND_LINEAR;
    if (recon5_applies(kind, M, N, D, dyadic_order, scheme == SKB_SCHEME_S1)) return 6;
    if (solver_rows_per_lane(M, dyadic_order) < 0) return 7;      // materialised grids (skb_generic_adj.cu): any length
    return adjoint5_applies(kind, M, N, D, dyadic_order, scheme == SKB_SCHEME_S1) ? 5 : 1;
}

int skb_adjoint_sym_supported(int M, int D, int dyadic_order, int static_kind, int scheme) {
    if (M < 2 || D <= 0 || dyadic_order < 0 || dyadic_order > 20) return 0;
    if (static_kind != SKB_STATIC_LINEAR && static_kind != SKB_STATIC_RBF) return 0;
    const int kind = static_kind == SKB_STATIC_RBF ? KIND_RBF : KIND_LINEAR;
    return recon5_applies(kind, M, M, D, dyadic_order, scheme == SKB_SCHEME_S1) &&
           recon5_sym_applies(kind, M, M, D, dyadic_order, scheme == SKB_SCHEME_S1) ? 1 : 0;
}

size_t skb_fwd_workspace_bytes(int A, int B, int M, int N, int D, int dyadic_order, int pairs) {
    if (A <= 0 || B <= 0 || M < 2 || N < 2 || D <= 0 || dyadic_order < 0 || dyadic_order > 20) return 0;
    const size_t Dp = (size_t)padded_dim(D);
    size_t w = kCounterBytes + align256((size_t)A * M * Dp * sizeof(double)) + align256((size_t)B * N * Dp * sizeof(double));
    // (sized for either static kind and scheme: the tile path's part is added whenever the shape could take it)
    if (tile_applies(KIND_RBF, A, B, M, N, D, dyadic_order, false, pairs))
        w += tile_workspace_bytes(A, B, M, N, dyadic_order, pairs);
    if (solver_rows_per_lane(M, dyadic_order) < 0)
        w += generic_workspace_bytes(njobs_of(A, B, pairs == SKB_PAIRS_BATCH ? SKB_PAIRS_BATCH : SKB_PAIRS_GRAM), M, N, dyadic_order, true);
    return w;
}

size_t skb_aux_workspace_bytes(int A, int B, int M, int N, int dyadic_order, int pairs) {
    if (A <= 0 || B <= 0 || M < 2 || N < 2 || dyadic_order < 0 || dyadic_order > 20) return 0;
    size_t w = kCounterBytes;
    if (solver_rows_per_lane(M, dyadic_order) < 0)
        w += generic_workspace_bytes(njobs_of(A, B, pairs == SKB_PAIRS_BATCH ? SKB_PAIRS_BATCH : SKB_PAIRS_GRAM), M, N, dyadic_order, false);
    return w;
}

static bool recon_ok(int static_kind, int A, int B, int M, int N, int D, int d, int scheme) {
    const int kind = static_kind == SKB_STATIC_RBF ? KIND_RBF : KIND_LINEAR;
    const size_t Dp = (size_t)padded_dim(D);
    return recon5_applies(kind, M, N, D, d, scheme == SKB_SCHEME_S1) && (size_t)A * M * Dp * sizeof(double) < ((size_t)1 << 32) &&
           (size_t)B * N * Dp * sizeof(double) < ((size_t)1 << 32);
}

size_t skb_ctx_bytes(int A, int B, int M, int N, int dyadic_order, int pairs) {
    if (A <= 0 || B <= 0 || M < 2 || N < 2 || dyadic_order < 0 || dyadic_order > 20) return 0;
    return align256((size_t)njobs_of(A, B, pairs) * (ctx_row_doubles(N, dyadic_order) + ctx_col_doubles(M, dyadic_order)) * sizeof(double));
}

// ---- backward on materialised grids (any shape; skb_generic_adj.cu) -----------------------------------------------
static const size_t kMaterializedBudget = (size_t)2 << 30;
static size_t materialized_doubles_per_pair(int M, int N, int d) {
    const size_t MMf = (size_t)(M - 1) << d, NNf = (size_t)(N - 1) << d;
    return (size_t)M * N + 2 * (size_t)(M - 1) * (N - 1) + 2 * (MMf + 1) * (NNf + 1);
}

// a.Xp / a.Yp prepared (forward orientation, plain prep_factors), a.gscale set; out (njobs), grad (njobs, M, D)
static int run_materialized_adjoint(int kind, const KArgs& a, int d, long njobs, double* out, double* grad, char* scratch,
                                    size_t scratch_bytes, cudaStream_t st) {
    const int M = a.M, N = a.N;
    if ((((long)(M - 1)) << d) > 0x3fffffffL || (((long)(N - 1)) << d) > 0x3fffffffL) return SKB_ERR_BAD_SHAPE;
    const size_t per = materialized_doubles_per_pair(M, N, d) * sizeof(double);
    if (scratch_bytes < per + 512) return SKB_ERR_WORKSPACE;
    long chunk = (long)((scratch_bytes - 512) / per);
    if (chunk > njobs) chunk = njobs;
    if (chunk > 0x3fffffffL) chunk = 0x3fffffffL;
    double* base = (double*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    const size_t grid = (size_t)((((long)(M - 1)) << d) + 1) * (size_t)((((long)(N - 1)) << d) + 1);
    for (long j0 = 0; j0 < njobs; j0 += chunk) {
        const long nj = njobs - j0 < chunk ? njobs - j0 : chunk;
        double* Ks = base;
        double* incc = Ks + (size_t)nj * M * N;
        double* S = incc + (size_t)nj * (M - 1) * (N - 1);
        double* U = S + (size_t)nj * (M - 1) * (N - 1);
        (void)grid;
        int rc = launch_static_matrix(a, kind, j0, nj, Ks, st);
        if (rc) return rc;
        rc = launch_coarse_increments(Ks, incc, nj, M, N, a.scale4, false, st);
        if (rc) return rc;
        rc = launch_grid_solve(incc, U, out, j0, nj, M, N, d, a.s1 != 0, st);
        if (rc) return rc;
        rc = launch_coarse_sens(U, S, nj, M, N, d, a.scale4, st);
        if (rc) return rc;
        rc = launch_grad_from_sens(S, Ks, a, kind, grad, j0, nj, st);
        if (rc) return rc;
    }
    return SKB_OK;
}

// plugin kernels: Ks (njobs, M, N) given; out (njobs), S (njobs, M-1, N-1)
static int run_materialized_sensitivity(const double* Ks, int M, int N, int d, bool s1, double scale4, long njobs, double* out, double* S,
                                        char* scratch, size_t scratch_bytes, cudaStream_t st) {
    if ((((long)(M - 1)) << d) > 0x3fffffffL || (((long)(N - 1)) << d) > 0x3fffffffL) return SKB_ERR_BAD_SHAPE;
    const size_t MMf = (size_t)(M - 1) << d, NNf = (size_t)(N - 1) << d;
    const size_t per = ((size_t)(M - 1) * (N - 1) + 2 * (MMf + 1) * (NNf + 1)) * sizeof(double);
    if (scratch_bytes < per + 512) return SKB_ERR_WORKSPACE;
    long chunk = (long)((scratch_bytes - 512) / per);
    if (chunk > njobs) chunk = njobs;
    double* base = (double*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    for (long j0 = 0; j0 < njobs; j0 += chunk) {
        const long nj = njobs - j0 < chunk ? njobs - j0 : chunk;
        double* incc = base;
        double* U = incc + (size_t)nj * (M - 1) * (N - 1);
        int rc = launch_coarse_increments(Ks + (size_t)j0 * M * N, incc, nj, M, N, scale4, false, st);
        if (rc) return rc;
        rc = launch_grid_solve(incc, U, out, j0, nj, M, N, d, s1, st);
        if (rc) return rc;
        rc = launch_coarse_sens(U, S + (size_t)j0 * (M - 1) * (N - 1), nj, M, N, d, scale4, st);
        if (rc) return rc;
    }
    return SKB_OK;
}

// fixed part: counter block, prepared paths in both orientations, [boundary context]; then the stored-grid scratch
static size_t bwd_workspace_bytes(int A, int B, int M, int N, int D, int d, int pairs, bool with_ctx, bool with_vjp) {
    if (A <= 0 || B <= 0 || M < 2 || N < 2 || D <= 0 || d < 0) return 0;
    const long nj = njobs_of(A, B, pairs == SKB_PAIRS_BATCH ? SKB_PAIRS_BATCH : SKB_PAIRS_GRAM);
    const bool recon = recon_ok(SKB_STATIC_RBF, A, B, M, N, D, d, SKB_SCHEME_S2);
    const size_t grid_b = grid_doubles_per_pair(M, N, d) * sizeof(double);
    const size_t Dp = (size_t)padded_dim(D);
    if (!recon && grid_b == 0) {
        // materialised-grid backward (skb_sigkernel_fwd_bwd only): prepared paths + a chunk of pairs
        if (with_vjp) return 0;
        const size_t per = materialized_doubles_per_pair(M, N, d) * sizeof(double);
        size_t jobs = (size_t)nj;
        size_t cap = kMaterializedBudget / per;
        if (cap < 1) cap = 1;
        if (jobs > cap) jobs = cap;
        return kCounterBytes + align256((size_t)A * M * Dp * sizeof(double)) + align256((size_t)B * N * Dp * sizeof(double)) + jobs * per + 1024;
    }
    size_t w = kCounterBytes + 2 * align256((size_t)A * M * Dp * sizeof(double)) + 2 * align256((size_t)B * N * Dp * sizeof(double));
    if (recon && with_ctx) w += skb_ctx_bytes(A, B, M, N, d, pairs == SKB_PAIRS_BATCH ? SKB_PAIRS_BATCH : SKB_PAIRS_GRAM);
    if (with_vjp) w += align256((size_t)nj * sizeof(double)) + align256((size_t)A * M * D * sizeof(double));   // k of the fallback's forward pass; this call's gradient
    if (grid_b != 0) {
        const size_t per = grid_b + (with_vjp ? (size_t)M * D * sizeof(double) : 0);
        size_t jobs = (size_t)nj;
        size_t cap = (recon ? kFallbackBudget : kScratchBudget) / per;
        if (cap < 1) cap = 1;
        if (jobs > cap) jobs = cap;
        w += align256(front_pad_doubles(M, d) * sizeof(double)) + align256(jobs * per) + 512;
    }
    return w;
}

size_t skb_bwd_workspace_bytes(int A, int B, int M, int N, int D, int dyadic_order, int pairs) {
    return bwd_workspace_bytes(A, B, M, N, D, dyadic_order, pairs, true, false);
}

size_t skb_bwd_vjp_workspace_bytes(int A, int B, int M, int N, int D, int dyadic_order, int pairs) {
    return bwd_workspace_bytes(A, B, M, N, D, dyadic_order, pairs, false, true);
}

static int sigkernel_fwd_impl(const void* X, const void* Y, int io_dtype, int A, int B, int M, int N, int D,
                              int dyadic_order, int static_kind, double static_param, int scheme, int pairs,
                              int arith, double* out, double* const* out_peers, int n_peers, void* workspace,
                              size_t workspace_bytes, void* stream, long job_lo = 0, long job_hi = -1,
                              unsigned long long* const* sig_peers = nullptr, int sig_rank = 0, unsigned long long sig_epoch = 0) {
    int rc = check_common(A, B, M, N, dyadic_order, scheme, pairs, arith);
    if (rc) return rc;
    if (D <= 0) return SKB_ERR_BAD_SHAPE;
    if (static_kind != SKB_STATIC_LINEAR && static_kind != SKB_STATIC_RBF) return SKB_ERR_BAD_ENUM;
    if (io_dtype != SKB_F64 && io_dtype != SKB_F32) return SKB_ERR_BAD_ENUM;
    if (arith != SKB_ARITH_FMA) return SKB_ERR_BAD_ENUM;
    if (!X || !Y || (!out && n_peers == 0)) return SKB_ERR_NULL;
    if (n_peers < 0 || n_peers > 8 || (n_peers > 0 && !out_peers)) return SKB_ERR_BAD_ENUM;
    const size_t fixed_bytes = kCounterBytes + align256((size_t)A * M * padded_dim(D) * sizeof(double)) +
                               align256((size_t)B * N * padded_dim(D) * sizeof(double));
    if (!workspace || workspace_bytes < fixed_bytes) return SKB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int Dp = padded_dim(D);
    char* w = (char*)workspace;
    unsigned int* counter = (unsigned int*)w;
    double* Xp = (double*)(w + kCounterBytes);
    double* Yp = (double*)(w + kCounterBytes + align256((size_t)A * M * Dp * sizeof(double)));
    double cx, nsc;
    prep_factors(static_kind, static_param, cx, nsc);
    const int kind = static_kind == SKB_STATIC_RBF ? KIND_RBF : KIND_LINEAR;
    const bool use_tile = n_peers == 0 && tile_applies(kind, A, B, M, N, D, dyadic_order, scheme == SKB_SCHEME_S1, pairs) &&
                          workspace_bytes >= fixed_bytes + tile_workspace_bytes(A, B, M, N, dyadic_order, pairs);
    const bool use5 = !use_tile && fwd5_applies(kind, M, N, D, dyadic_order, scheme == SKB_SCHEME_S1) &&
                      (size_t)A * M * padded_dim(D) * sizeof(double) < ((size_t)1 << 32) &&
                      (size_t)B * N * padded_dim(D) * sizeof(double) < ((size_t)1 << 32);   // 32-bit byte offsets in the job ring
    if ((use5 || use_tile) && kind == KIND_LINEAR) cx *= fwd5_kscale(dyadic_order);   // k is produced pre-scaled on those paths
    if (((use5 && fwd5_scaled_exp(M, dyadic_order, D)) || use_tile) && kind == KIND_RBF) { cx *= tile_arg_scale(); nsc *= tile_arg_scale(); }   // exp argument in units of ln2 / 2048
    rc = launch_prep2(X, Y, io_dtype, Xp, nullptr, Yp, nullptr, A, M, B, N, D, Dp, cx, nsc, counter, st);
    if (rc) return rc;

    KArgs a = base_args(A, B, M, N, dyadic_order, scheme, pairs);
    a.Xp = Xp; a.Yp = Yp; a.out = out; a.counter = counter;
    a.counter_clean = 1;                 // zeroed by the preparation kernel
    const long nj = njobs_of(A, B, pairs);
    if (nj > 0x7fffffffL) return SKB_ERR_BAD_SHAPE;
    a.njobs = (int)nj;
    a.Dp = Dp; a.D = D;
    const bool ranged = job_hi >= 0;
    if (ranged) {
        // a slice [job_lo, job_hi) of the pair enumeration (one rank's share of a sharded Gram matrix)
        if (job_lo < 0 || job_hi < job_lo || job_hi > nj) return SKB_ERR_BAD_SHAPE;
        if (!use5) return SKB_ERR_UNSUPPORTED;
        if (sig_peers) {
            if (n_peers <= 0 || sig_rank < 0 || sig_rank >= n_peers || sig_epoch == 0) return SKB_ERR_BAD_ENUM;
            for (int q = 0; q < n_peers; ++q) {
                if (!sig_peers[q]) return SKB_ERR_NULL;
                a.sig_peer[q] = sig_peers[q];
            }
            a.sig_rank = sig_rank;
            a.sig_epoch = sig_epoch;
            a.n_peer = n_peers;
        }
        if (job_hi == job_lo) return sig_peers ? launch_rank_barrier(a, st) : SKB_OK;     // nothing to solve: the barrier alone
        a.job0 = job_lo;
        a.njobs = (int)(job_hi - job_lo);
    }
    if (n_peers > 0) {
        // results go straight to every rank's copy of G: only the fwd5 kernels have that output path
        if (!use5) return SKB_ERR_UNSUPPORTED;
        for (int q = 0; q < n_peers; ++q) {
            if (!out_peers[q]) return SKB_ERR_NULL;
            a.out_peer[q] = out_peers[q];
        }
        a.n_peer = n_peers;
        return launch_forward5(kind, dyadic_order, a, st);
    }
    if (use_tile) return launch_tile_forward(kind, dyadic_order, a, w + fixed_bytes, st);
    if (use5) return launch_forward5(kind, dyadic_order, a, st);
    if (solver_rows_per_lane(M, dyadic_order) >= 0) return launch_solver(MODE_FWD, kind, dyadic_order, false, a, st);
    // shape outside the register-resident kernels: generic row-band fallback (a symmetric request is
    // served by solving the full square)
    if (pairs == SKB_PAIRS_SYM) a.pairs = SKB_PAIRS_GRAM;
    return run_generic_forward(kind, a, dyadic_order, false, njobs_of(A, B, a.pairs), w + fixed_bytes,
                               workspace_bytes - fixed_bytes, st);
}

int skb_sigkernel_fwd(const void* X, const void* Y, int io_dtype, int A, int B, int M, int N, int D,
                      int dyadic_order, int static_kind, double static_param, int scheme, int pairs,
                      int arith, double* out, void* workspace, size_t workspace_bytes, void* stream) {
    return sigkernel_fwd_impl(X, Y, io_dtype, A, B, M, N, D, dyadic_order, static_kind, static_param, scheme, pairs, arith, out,
                              nullptr, 0, workspace, workspace_bytes, stream);
}

int skb_sigkernel_fwd_peers(const void* X, const void* Y, int io_dtype, int A, int B, int M, int N, int D,
                            int dyadic_order, int static_kind, double static_param, int scheme, int pairs,
                            double* const* out_peers, int n_peers, void* workspace, size_t workspace_bytes, void* stream) {
    if (n_peers <= 0) return SKB_ERR_BAD_ENUM;
    return sigkernel_fwd_impl(X, Y, io_dtype, A, B, M, N, D, dyadic_order, static_kind, static_param, scheme, pairs, SKB_ARITH_FMA,
                              nullptr, out_peers, n_peers, workspace, workspace_bytes, stream);
}

int skb_sigkernel_fwd_range(const void* X, const void* Y, int io_dtype, int A, int B, int M, int N, int D,
                            int dyadic_order, int static_kind, double static_param, int scheme, int pairs,
                            long job_lo, long job_hi, double* out, double* const* out_peers, int n_peers,
                            unsigned long long* const* sig_peers, int my_rank, unsigned long long sig_epoch,
                            void* workspace, size_t workspace_bytes, void* stream) {
    if (n_peers < 0 || job_hi < 0) return SKB_ERR_BAD_ENUM;
    return sigkernel_fwd_impl(X, Y, io_dtype, A, B, M, N, D, dyadic_order, static_kind, static_param, scheme, pairs, SKB_ARITH_FMA,
                              out, out_peers, n_peers, workspace, workspace_bytes, stream, job_lo, job_hi, sig_peers, my_rank, sig_epoch);
}

int skb_static_gram(const void* X, const void* Y, int io_dtype, int A, int B, int M, int N, int D, int static_kind,
                    double static_param, int pairs, double* Ks, void* workspace, size_t workspace_bytes, void* stream) {
    if (A <= 0 || B <= 0 || M < 1 || N < 1 || D <= 0) return SKB_ERR_BAD_SHAPE;
    if (pairs != SKB_PAIRS_GRAM && pairs != SKB_PAIRS_BATCH) return SKB_ERR_BAD_ENUM;
    if (pairs == SKB_PAIRS_BATCH && A != B) return SKB_ERR_BAD_SHAPE;
    if (static_kind != SKB_STATIC_LINEAR && static_kind != SKB_STATIC_RBF) return SKB_ERR_BAD_ENUM;
    if (io_dtype != SKB_F64 && io_dtype != SKB_F32) return SKB_ERR_BAD_ENUM;
    if (!X || !Y || !Ks) return SKB_ERR_NULL;
    const int Dp = padded_dim(D);
    const size_t xb = align256((size_t)A * M * Dp * sizeof(double)), yb = align256((size_t)B * N * Dp * sizeof(double));
    if (!workspace || workspace_bytes < kCounterBytes + xb + yb) return SKB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)workspace;
    double* Xp = (double*)(w + kCounterBytes);
    double* Yp = (double*)(w + kCounterBytes + xb);
    double cx, nsc;
    prep_factors(static_kind, static_param, cx, nsc);
    int rc = launch_prep2(X, Y, io_dtype, Xp, nullptr, Yp, nullptr, A, M, B, N, D, Dp, cx, nsc, nullptr, st);
    if (rc) return rc;
    KArgs a = base_args(A, B, M, N, 0, SKB_SCHEME_S2, pairs);
    a.Xp = Xp; a.Yp = Yp; a.Dp = Dp; a.D = D;
    return launch_static_matrix(a, static_kind == SKB_STATIC_RBF ? KIND_RBF : KIND_LINEAR, 0, njobs_of(A, B, pairs), Ks, st);
}

int skb_sigkernel_fwd_from_static(const double* Ks, int A, int B, int M, int N, int dyadic_order, int scheme,
                                  int pairs, int arith, double* out, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    int rc = check_common(A, B, M, N, dyadic_order, scheme, pairs, arith);
    if (rc) return rc;
    if (!Ks || !out) return SKB_ERR_NULL;
    if (!workspace || workspace_bytes < kCounterBytes) return SKB_ERR_WORKSPACE;
    KArgs a = base_args(A, B, M, N, dyadic_order, scheme, pairs);
    a.Ks = Ks; a.out = out; a.counter = (unsigned int*)workspace;
    const long nj = njobs_of(A, B, pairs);
    if (nj > 0x7fffffffL) return SKB_ERR_BAD_SHAPE;
    a.njobs = (int)nj;
    if (solver_rows_per_lane(M, dyadic_order) >= 0)
        return launch_solver(MODE_FWD, KIND_STATIC, dyadic_order, arith == SKB_ARITH_EXACT, a, (cudaStream_t)stream);
    if (pairs == SKB_PAIRS_SYM) a.pairs = SKB_PAIRS_GRAM;
    return run_generic_forward(KIND_STATIC, a, dyadic_order, arith == SKB_ARITH_EXACT, njobs_of(A, B, a.pairs),
                               (char*)workspace + kCounterBytes, workspace_bytes - kCounterBytes, (cudaStream_t)stream);
}

int skb_sigkernel_solve_increments(const double* inc, long P, int MM, int NN, int scheme, int arith,
                                   double* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (P <= 0 || MM < 1 || NN < 1 || P > 0x7fffffffL) return SKB_ERR_BAD_SHAPE;
    if (scheme != SKB_SCHEME_S2 && scheme != SKB_SCHEME_S1) return SKB_ERR_BAD_ENUM;
    if (arith != SKB_ARITH_FMA && arith != SKB_ARITH_EXACT) return SKB_ERR_BAD_ENUM;
    if (!inc || !out) return SKB_ERR_NULL;
    if (!workspace || workspace_bytes < kCounterBytes) return SKB_ERR_WORKSPACE;
    KArgs a = base_args((int)P, (int)P, MM + 1, NN + 1, 0, scheme, SKB_PAIRS_BATCH);
    a.Mv = MM; a.Nv = NN;
    a.Ks = inc; a.out = out; a.counter = (unsigned int*)workspace;
    a.njobs = (int)P;
    if (solver_rows_per_lane(MM + 1, 0) >= 0)
        return launch_solver(MODE_FWD, KIND_INC, 0, arith == SKB_ARITH_EXACT, a, (cudaStream_t)stream);
    return run_generic_forward(KIND_INC, a, 0, arith == SKB_ARITH_EXACT, P, (char*)workspace + kCounterBytes,
                               workspace_bytes - kCounterBytes, (cudaStream_t)stream);
}

// fused loss head on the stored-grid path: per-point gradients of a chunk go to a chunk-sized buffer and are contracted
// with d loss / d K right after the chunk's reversed sweep
struct VjpOpts {
    const double* gout;
    double w_diag, w_off;
    double* gradX;
};

// shared driver of the backward entry points: forward-with-store then reversed sweep, in chunks of pairs whose forward
// grids (and, with a loss head, per-point gradients) fit the scratch part of the workspace.  fa.cond / ra.cond (if set)
// make every launch a no-op unless the flag they point to is raised.
static int run_adjoint(int kind, int rev_mode, KArgs fa, KArgs ra, int d, long njobs, double* scratch_base,
                       size_t scratch_bytes, cudaStream_t st, bool v5 = false, const VjpOpts* vjp = nullptr) {
    const size_t grid_b = grid_doubles_per_pair(fa.M, fa.N, d) * sizeof(double);
    const size_t gp_b = vjp ? (size_t)fa.M * fa.D * sizeof(double) : 0;
    const size_t per = grid_b + gp_b;
    const size_t pad = front_pad_doubles(fa.M, d) * sizeof(double);
    if (grid_b == 0) return SKB_ERR_UNSUPPORTED;
    if (scratch_bytes < align256(pad) + per + 256) return SKB_ERR_WORKSPACE;
    long chunk = (long)((scratch_bytes - align256(pad) - 256) / per);
    if (chunk > njobs) chunk = njobs;
    double* grid = (double*)((char*)scratch_base + align256(pad));
    double* gpbuf = vjp ? (double*)((char*)grid + align256((size_t)chunk * grid_b)) : nullptr;
    if (vjp && (char*)gpbuf + (size_t)chunk * gp_b > (char*)scratch_base + scratch_bytes) --chunk;
    if (chunk < 1) return SKB_ERR_WORKSPACE;
    for (long j0 = 0; j0 < njobs; j0 += chunk) {
        const int nj = (int)(njobs - j0 < chunk ? njobs - j0 : chunk);
        fa.job0 = ra.job0 = j0;
        fa.njobs = ra.njobs = nj;
        fa.scratch = ra.scratch = grid;
        fa.counter_clean = ra.counter_clean = 0;
        if (vjp) ra.grad = gpbuf - (size_t)j0 * fa.M * fa.D;      // the kernels index per-point gradients by the global pair index
        int rc = v5 ? launch_adjoint5(MODE_FWD_STORE, kind, d, fa, st) : launch_solver(MODE_FWD_STORE, kind, d, false, fa, st);
        if (rc) return rc;
        rc = v5 ? launch_adjoint5(rev_mode, kind, d, ra, st) : launch_solver(rev_mode, kind, d, false, ra, st);
        if (rc) return rc;
        if (vjp) {
            rc = launch_vjp_accumulate(gpbuf, j0, nj, fa.A, fa.B, fa.M, fa.D, fa.pairs, vjp->gout, vjp->w_diag, vjp->w_off,
                                       vjp->gradX, ra.cond, st);
            if (rc) return rc;
        }
    }
    return SKB_OK;
}

int skb_sigkernel_fwd_bwd(const void* X, const void* Y, int io_dtype, int A, int B, int M, int N, int D,
                          int dyadic_order, int static_kind, double static_param, int scheme, int pairs,
                          double* out, double* grad_points, void* workspace, size_t workspace_bytes,
                          void* stream) {
    int rc = check_common(A, B, M, N, dyadic_order, scheme, pairs, SKB_ARITH_FMA);
    if (rc) return rc;
    if (D <= 0) return SKB_ERR_BAD_SHAPE;
    if (static_kind != SKB_STATIC_LINEAR && static_kind != SKB_STATIC_RBF) return SKB_ERR_BAD_ENUM;
    if (io_dtype != SKB_F64 && io_dtype != SKB_F32) return SKB_ERR_BAD_ENUM;
    if (!X || !Y || !out || !grad_points) return SKB_ERR_NULL;
    if (!workspace) return SKB_ERR_WORKSPACE;
    // SYM (Y = X): out and grad_points are the full (A, A) / (A, A, M, D) tensors; one forward solve and one reversed sweep per
    // unordered pair where the unordered-pair sweep covers the shape, the full square otherwise (the same numbers)
    const bool sym_sweep = pairs == SKB_PAIRS_SYM && recon_ok(static_kind, A, B, M, N, D, dyadic_order, scheme) &&
                           recon5_sym_applies(static_kind == SKB_STATIC_RBF ? KIND_RBF : KIND_LINEAR, M, N, D, dyadic_order, scheme == SKB_SCHEME_S1);
    if (pairs == SKB_PAIRS_SYM && !sym_sweep)
        return skb_sigkernel_fwd_bwd(X, Y, io_dtype, A, B, M, N, D, dyadic_order, static_kind, static_param, scheme, SKB_PAIRS_GRAM, out,
                                     grad_points, workspace, workspace_bytes, stream);
    cudaStream_t st = (cudaStream_t)stream;
    const int Dp = padded_dim(D);
    const size_t xb = align256((size_t)A * M * Dp * sizeof(double)), yb = align256((size_t)B * N * Dp * sizeof(double));
    const long nj = njobs_of(A, B, pairs);
    if (nj > 0x7fffffffL) return SKB_ERR_BAD_SHAPE;
    const bool recon = recon_ok(static_kind, A, B, M, N, D, dyadic_order, scheme);
    const bool stored = solver_rows_per_lane(M, dyadic_order) >= 0;
    if (!recon && !stored) {
        // outside every register-resident adjoint kernel: the reference's algebra on materialised grids
        const int Dpm = padded_dim(D);
        const size_t xbm = align256((size_t)A * M * Dpm * sizeof(double)), ybm = align256((size_t)B * N * Dpm * sizeof(double));
        if (workspace_bytes < kCounterBytes + xbm + ybm) return SKB_ERR_WORKSPACE;
        char* wm = (char*)workspace;
        double* Xm = (double*)(wm + kCounterBytes);
        double* Ym = (double*)(wm + kCounterBytes + xbm);
        double cxm, nscm;
        prep_factors(static_kind, static_param, cxm, nscm);
        rc = launch_prep2(X, Y, io_dtype, Xm, nullptr, Ym, nullptr, A, M, B, N, D, Dpm, cxm, nscm, (unsigned int*)wm, st);
        if (rc) return rc;
        KArgs ma = base_args(A, B, M, N, dyadic_order, scheme, pairs);
        ma.Xp = Xm; ma.Yp = Ym; ma.Dp = Dpm; ma.D = D;
        ma.gscale = static_kind == SKB_STATIC_RBF ? 2.0 / static_param : static_param;
        return run_materialized_adjoint(static_kind == SKB_STATIC_RBF ? KIND_RBF : KIND_LINEAR, ma, dyadic_order, nj, out, grad_points,
                                        wm + kCounterBytes + xbm + ybm, workspace_bytes - (kCounterBytes + xbm + ybm), st);
    }
    const size_t ctxb = recon ? skb_ctx_bytes(A, B, M, N, dyadic_order, pairs) : 0;
    const size_t fixed = kCounterBytes + 2 * xb + 2 * yb + ctxb;
    if (workspace_bytes < fixed) return SKB_ERR_WORKSPACE;
    char* w = (char*)workspace;
    unsigned int* counter = (unsigned int*)w;
    unsigned int* flag = (unsigned int*)(w + kFlagOffset);
    double* Xp = (double*)(w + kCounterBytes);
    double* Yp = (double*)(w + kCounterBytes + xb);
    double* Xr = (double*)(w + kCounterBytes + xb + yb);
    double* Yr = (double*)(w + kCounterBytes + 2 * xb + yb);
    void* ctx = w + kCounterBytes + 2 * xb + 2 * yb;
    double cx, nsc;
    prep_factors(static_kind, static_param, cx, nsc);
    const int kind5 = static_kind == SKB_STATIC_RBF ? KIND_RBF : KIND_LINEAR;
    const bool v5 = adjoint5_applies(kind5, M, N, D, dyadic_order, scheme == SKB_SCHEME_S1) &&
                    (size_t)A * M * Dp * sizeof(double) < ((size_t)1 << 32) && (size_t)B * N * Dp * sizeof(double) < ((size_t)1 << 32);
    // (the stored-grid fallback of a reconstruction call must read the same prepared rows: it then needs the v5 kernels)
    const bool fallback = recon && stored && (kind5 != KIND_LINEAR || v5) &&
                          workspace_bytes >= fixed + align256(front_pad_doubles(M, dyadic_order) * sizeof(double)) +
                                                 grid_doubles_per_pair(M, N, dyadic_order) * sizeof(double) + 512;
    if ((recon || v5) && kind5 == KIND_LINEAR) cx *= fwd5_kscale(dyadic_order);   // k is produced pre-scaled on the v5 paths
    rc = check_cuda(cudaMemsetAsync(w, 0, kCounterBytes, st));
    if (rc) return rc;
    rc = launch_prep2(X, Y, io_dtype, Xp, Xr, Yp, Yr, A, M, B, N, D, Dp, cx, nsc, nullptr, st);
    if (rc) return rc;

    KArgs fa = base_args(A, B, M, N, dyadic_order, scheme, pairs);
    fa.Xp = Xp; fa.Yp = Yp; fa.out = out; fa.counter = counter; fa.Dp = Dp; fa.D = D;
    fa.njobs = (int)nj;
    KArgs ra = fa;
    ra.Xp = Xr; ra.Yp = Yr; ra.out = nullptr; ra.grad = grad_points;
    ra.gscale = static_kind == SKB_STATIC_RBF ? 2.0 / static_param : static_param;
    if (recon) {
        set_ctx(fa, ctx, nj, M, N, dyadic_order);
        set_ctx(ra, ctx, nj, M, N, dyadic_order);
        fa.counter_clean = 1;
        rc = launch_recon5(MODE_FWD_EMIT, kind5, dyadic_order, fa, st);
        if (rc) return rc;
        ra.flag = flag; ra.recon_tol = kReconTol;
        rc = launch_recon5(sym_sweep ? MODE_REV_RECON_SYM : MODE_REV_RECON, kind5, dyadic_order, ra, st);
        if (rc) return rc;
        if (!fallback) return SKB_OK;
        fa.cond = ra.cond = flag;            // queued behind a device-side test of the flag
        if (sym_sweep) {
            // the stored-grid kernels run over the ordered pairs of the full square (same out, same grad_points)
            fa.pairs = ra.pairs = SKB_PAIRS_GRAM;
            return run_adjoint(kind5, MODE_REV_GRAD, fa, ra, dyadic_order, njobs_of(A, B, SKB_PAIRS_GRAM), (double*)(w + fixed),
                               workspace_bytes - fixed, st, v5);
        }
    }
    return run_adjoint(kind5, MODE_REV_GRAD, fa, ra, dyadic_order, nj, (double*)(w + fixed), workspace_bytes - fixed, st, v5);
}

int skb_sigkernel_fwd_ctx(const void* X, const void* Y, int io_dtype, int A, int B, int M, int N, int D, int dyadic_order,
                          int static_kind, double static_param, int scheme, int pairs, double* out, void* ctx,
                          size_t ctx_bytes, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(A, B, M, N, dyadic_order, scheme, pairs, SKB_ARITH_FMA);
    if (rc) return rc;
    if (D <= 0) return SKB_ERR_BAD_SHAPE;
    if (static_kind != SKB_STATIC_LINEAR && static_kind != SKB_STATIC_RBF) return SKB_ERR_BAD_ENUM;
    if (io_dtype != SKB_F64 && io_dtype != SKB_F32) return SKB_ERR_BAD_ENUM;
    if (!X || !Y || !out || !ctx) return SKB_ERR_NULL;
    if (!recon_ok(static_kind, A, B, M, N, D, dyadic_order, scheme)) return SKB_ERR_UNSUPPORTED;
    if (ctx_bytes < skb_ctx_bytes(A, B, M, N, dyadic_order, pairs)) return SKB_ERR_WORKSPACE;
    const int Dp = padded_dim(D);
    const size_t xb = align256((size_t)A * M * Dp * sizeof(double)), yb = align256((size_t)B * N * Dp * sizeof(double));
    if (!workspace || workspace_bytes < kCounterBytes + xb + yb) return SKB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)workspace;
    unsigned int* counter = (unsigned int*)w;
    double* Xp = (double*)(w + kCounterBytes);
    double* Yp = (double*)(w + kCounterBytes + xb);
    double cx, nsc;
    prep_factors(static_kind, static_param, cx, nsc);
    const int kind5 = static_kind == SKB_STATIC_RBF ? KIND_RBF : KIND_LINEAR;
    if (kind5 == KIND_LINEAR) cx *= fwd5_kscale(dyadic_order);
    rc = launch_prep2(X, Y, io_dtype, Xp, nullptr, Yp, nullptr, A, M, B, N, D, Dp, cx, nsc, counter, st);
    if (rc) return rc;
    const long nj = njobs_of(A, B, pairs);
    if (nj > 0x7fffffffL) return SKB_ERR_BAD_SHAPE;
    KArgs fa = base_args(A, B, M, N, dyadic_order, scheme, pairs);
    fa.Xp = Xp; fa.Yp = Yp; fa.out = out; fa.counter = counter; fa.Dp = Dp; fa.D = D;
    fa.njobs = (int)nj;
    fa.counter_clean = 1;
    set_ctx(fa, ctx, nj, M, N, dyadic_order);
    return launch_recon5(MODE_FWD_EMIT, kind5, dyadic_order, fa, st);
}

int skb_sigkernel_bwd_vjp(const void* X, const void* Y, int io_dtype, int A, int B, int M, int N, int D, int dyadic_order,
                          int static_kind, double static_param, int scheme, int pairs, const void* ctx, int ctx_pairs,
                          const double* grad_out, double w_diag, double w_off, double out_scale, const double* out_scale_dev,
                          int accumulate, double* gradX, double* grad_points,
                          void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(A, B, M, N, dyadic_order, scheme, pairs, SKB_ARITH_FMA);
    if (rc) return rc;
    // pairs = SYM (Gram(X, X), Y = X): ONE reversed sweep per unordered pair a <= b yields both d k / d X_a and d k / d X_b
    // (MODE_REV_RECON_SYM), i.e. the terms of the ordered pairs (a, b) and (b, a); no per-pair output
    const bool usym = pairs == SKB_PAIRS_SYM;
    if (usym && (ctx_pairs != SKB_PAIRS_SYM || A != B || M != N || grad_points || !gradX)) return SKB_ERR_BAD_ENUM;
    if (ctx_pairs != pairs && !(ctx_pairs == SKB_PAIRS_SYM && pairs == SKB_PAIRS_GRAM && A == B && M == N)) return SKB_ERR_BAD_ENUM;
    if (D <= 0) return SKB_ERR_BAD_SHAPE;
    if (static_kind != SKB_STATIC_LINEAR && static_kind != SKB_STATIC_RBF) return SKB_ERR_BAD_ENUM;
    if (io_dtype != SKB_F64 && io_dtype != SKB_F32) return SKB_ERR_BAD_ENUM;
    if (!X || !Y || !ctx || (!gradX && !grad_points)) return SKB_ERR_NULL;
    if (!recon_ok(static_kind, A, B, M, N, D, dyadic_order, scheme)) return SKB_ERR_UNSUPPORTED;
    if (usym && !recon5_sym_applies(static_kind == SKB_STATIC_RBF ? KIND_RBF : KIND_LINEAR, M, N, D, dyadic_order, scheme == SKB_SCHEME_S1))
        return SKB_ERR_UNSUPPORTED;
    const int Dp = padded_dim(D);
    const size_t xb = align256((size_t)A * M * Dp * sizeof(double)), yb = align256((size_t)B * N * Dp * sizeof(double));
    const long nj = njobs_of(A, B, usym ? SKB_PAIRS_GRAM : pairs);      // (sizes the fallback's part of the workspace)
    if (nj > 0x7fffffffL) return SKB_ERR_BAD_SHAPE;
    const size_t kb = align256((size_t)nj * sizeof(double));
    const size_t gb = align256((size_t)A * M * D * sizeof(double));       // gradient of this call before scaling / accumulation
    const size_t fixed = kCounterBytes + 2 * xb + 2 * yb + gb;
    if (!workspace || workspace_bytes < fixed) return SKB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)workspace;
    unsigned int* counter = (unsigned int*)w;
    unsigned int* flag = (unsigned int*)(w + kFlagOffset);
    double* Xp = (double*)(w + kCounterBytes);
    double* Yp = (double*)(w + kCounterBytes + xb);
    double* Xr = (double*)(w + kCounterBytes + xb + yb);
    double* Yr = (double*)(w + kCounterBytes + 2 * xb + yb);
    double cx, nsc;
    prep_factors(static_kind, static_param, cx, nsc);
    const int kind5 = static_kind == SKB_STATIC_RBF ? KIND_RBF : KIND_LINEAR;
    if (kind5 == KIND_LINEAR) cx *= fwd5_kscale(dyadic_order);
    const bool v5 = adjoint5_applies(kind5, M, N, D, dyadic_order, scheme == SKB_SCHEME_S1);
    const bool stored = solver_rows_per_lane(M, dyadic_order) >= 0 && (kind5 != KIND_LINEAR || v5);
    const size_t per = grid_doubles_per_pair(M, N, dyadic_order) * sizeof(double) + (size_t)M * D * sizeof(double);
    const bool fallback = stored && workspace_bytes >= fixed + kb + align256(front_pad_doubles(M, dyadic_order) * sizeof(double)) + per + 768;
    rc = check_cuda(cudaMemsetAsync(w, 0, kCounterBytes, st));
    if (rc) return rc;
    double* gtmp = (double*)(w + kCounterBytes + 2 * xb + 2 * yb);
    if (gradX) {
        rc = check_cuda(cudaMemsetAsync(gtmp, 0, (size_t)A * M * D * sizeof(double), st));
        if (rc) return rc;
    }
    rc = launch_prep2(X, Y, io_dtype, Xp, Xr, Yp, Yr, A, M, B, N, D, Dp, cx, nsc, nullptr, st);
    if (rc) return rc;
    KArgs ra = base_args(A, B, M, N, dyadic_order, scheme, pairs);
    ra.Xp = Xr; ra.Yp = Yr; ra.counter = counter; ra.Dp = Dp; ra.D = D;
    ra.njobs = (int)njobs_of(A, B, pairs);
    ra.counter_clean = 1;
    ra.grad = grad_points;
    ra.gscale = static_kind == SKB_STATIC_RBF ? 2.0 / static_param : static_param;
    set_ctx(ra, const_cast<void*>(ctx), njobs_of(A, B, ctx_pairs), M, N, dyadic_order);
    ra.bsym = ctx_pairs == SKB_PAIRS_SYM && pairs != SKB_PAIRS_SYM;
    ra.flag = flag; ra.recon_tol = kReconTol;
    ra.gout = grad_out; ra.gradX = gradX ? gtmp : nullptr; ra.w_diag = w_diag; ra.w_off = w_off;
    rc = launch_recon5(usym ? MODE_REV_RECON_SYM : MODE_REV_RECON, kind5, dyadic_order, ra, st);
    if (rc) return rc;
    if (!fallback) return gradX ? launch_combine(gradX, gtmp, (size_t)A * M * D, out_scale, out_scale_dev, accumulate, st) : SKB_OK;
    // stored-grid fallback behind a device-side test of the flag: forward with store, reversed sweep, loss head
    // (unordered pairs: the fallback runs over the full square, one sweep per ordered pair -- the same sum)
    KArgs fa = base_args(A, B, M, N, dyadic_order, scheme, usym ? SKB_PAIRS_GRAM : pairs);
    fa.Xp = Xp; fa.Yp = Yp; fa.out = (double*)(w + fixed); fa.counter = counter; fa.Dp = Dp; fa.D = D;
    KArgs rb = fa;
    rb.Xp = Xr; rb.Yp = Yr; rb.out = nullptr; rb.grad = grad_points;
    rb.gscale = ra.gscale;
    fa.cond = rb.cond = flag;
    VjpOpts vo = {grad_out, w_diag, w_off, gtmp};
    if (gradX) {
        rc = launch_cond_zero(gtmp, (size_t)A * M * D, flag, st);
        if (rc) return rc;
    }
    // (with grad_points requested the fallback writes them in place and the loss head reads them back chunk by chunk)
    rc = run_adjoint(kind5, MODE_REV_GRAD, fa, rb, dyadic_order, nj, (double*)(w + fixed + kb), workspace_bytes - fixed - kb, st, v5,
                     gradX ? &vo : nullptr);
    if (rc) return rc;
    return gradX ? launch_combine(gradX, gtmp, (size_t)A * M * D, out_scale, out_scale_dev, accumulate, st) : SKB_OK;
}

int skb_gram_weighted_sum(const double* G, int A, int B, int pairs, double w_diag, double w_off, double* acc, int accumulate,
                          void* stream) {
    if (A <= 0 || B <= 0) return SKB_ERR_BAD_SHAPE;
    if (pairs != SKB_PAIRS_GRAM && pairs != SKB_PAIRS_BATCH && pairs != SKB_PAIRS_SYM) return SKB_ERR_BAD_ENUM;
    if (!G || !acc) return SKB_ERR_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) {
        int rc = check_cuda(cudaMemsetAsync(acc, 0, sizeof(double), st));
        if (rc) return rc;
    }
    const long n = pairs == SKB_PAIRS_BATCH ? (long)A : (long)A * B;
    long blocks = (n + 255) / 256;
    if (blocks > 64) blocks = 64;
    gram_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(G, n, pairs == SKB_PAIRS_BATCH ? 0 : B, w_diag, w_off, acc);
    return check_launch();
}

size_t skb_sensitivity_workspace_bytes(int A, int B, int M, int N, int dyadic_order, int pairs) {
    if (A <= 0 || B <= 0 || M < 2 || N < 2 || dyadic_order < 0 || dyadic_order > 20) return 0;
    if (solver_rows_per_lane(M, dyadic_order) >= 0) return bwd_workspace_bytes(A, B, M, N, 1, dyadic_order, pairs, true, false);
    // materialised grids: coarse increments + the two fine grids of a chunk of pairs
    const size_t MMf = (size_t)(M - 1) << dyadic_order, NNf = (size_t)(N - 1) << dyadic_order;
    const size_t per = ((size_t)(M - 1) * (N - 1) + 2 * (MMf + 1) * (NNf + 1)) * sizeof(double);
    size_t jobs = (size_t)njobs_of(A, B, pairs == SKB_PAIRS_BATCH ? SKB_PAIRS_BATCH : SKB_PAIRS_GRAM);
    size_t cap = kMaterializedBudget / per;
    if (cap < 1) cap = 1;
    if (jobs > cap) jobs = cap;
    return kCounterBytes + jobs * per + 1024;
}

int skb_sigkernel_sensitivity_from_static(const double* Ks, int A, int B, int M, int N, int dyadic_order,
                                          int scheme, int pairs, double* out, double* S, void* workspace,
                                          size_t workspace_bytes, void* stream) {
    int rc = check_common(A, B, M, N, dyadic_order, scheme, pairs, SKB_ARITH_FMA);
    if (rc) return rc;
    if (pairs == SKB_PAIRS_SYM) return SKB_ERR_BAD_ENUM;
    if (!Ks || !out || !S) return SKB_ERR_NULL;
    if (!workspace || workspace_bytes < kCounterBytes) return SKB_ERR_WORKSPACE;
    const long nj = njobs_of(A, B, pairs);
    if (nj > 0x7fffffffL) return SKB_ERR_BAD_SHAPE;
    if (solver_rows_per_lane(M, dyadic_order) < 0)       // beyond the register-resident kernels: materialised grids, any length
        return run_materialized_sensitivity(Ks, M, N, dyadic_order, scheme == SKB_SCHEME_S1, scale4_of(dyadic_order), nj, out, S,
                                            (char*)workspace + kCounterBytes, workspace_bytes - kCounterBytes, (cudaStream_t)stream);
    KArgs fa = base_args(A, B, M, N, dyadic_order, scheme, pairs);
    fa.Ks = Ks; fa.out = out; fa.counter = (unsigned int*)workspace;
    KArgs ra = fa;
    ra.out = nullptr; ra.S = S;
    return run_adjoint(KIND_STATIC, MODE_REV_S, fa, ra, dyadic_order, nj, (double*)((char*)workspace + kCounterBytes),
                       workspace_bytes - kCounterBytes, (cudaStream_t)stream);
}

size_t skb_deriv_workspace_bytes(int A, int B, int M, int N) {
    if (A <= 0 || B <= 0 || M < 2 || N < 2) return 0;
    return align256((size_t)A * B * (M - 1) * (N - 1) * 3 * sizeof(double)) + 256;
}

int skb_sigkernel_derivatives_from_static(const double* K0, const double* K1, const double* K2, int A, int B, int M,
                                          int N, int dyadic_order, double eps, double* out3, void* workspace,
                                          size_t workspace_bytes, void* stream) {
    if (A <= 0 || B <= 0 || M < 2 || N < 2 || dyadic_order < 0 || dyadic_order > 20 || !(eps > 0.0)) return SKB_ERR_BAD_SHAPE;
    if (!K0 || !K1 || !K2 || !out3) return SKB_ERR_NULL;
    if (!workspace || workspace_bytes < skb_deriv_workspace_bytes(A, B, M, N)) return SKB_ERR_WORKSPACE;
    double* inc3 = (double*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    return launch_derivatives(K0, K1, K2, (long)A * B, M, N, dyadic_order, eps, inc3, out3, (cudaStream_t)stream);
}

}  // extern "C"
