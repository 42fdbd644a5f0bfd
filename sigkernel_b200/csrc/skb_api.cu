// skb_api.cu -- the extern "C" boundary declared in include/sigkernel_b200.h.
// Argument validation, workspace carving and kernel dispatch only; no allocation, no sync.
#include <string.h>
#include "skb_host.h"

namespace skb {

static thread_local int g_last_cuda = 0;

int check_cuda(cudaError_t e) {
    if (e == cudaSuccess) return SKB_OK;
    g_last_cuda = (int)e;
    return SKB_ERR_CUDA;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline int padded_dim(int D) { return (D + 2) & ~1; }   // 1 norm slot + D, rounded up to even

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

}  // namespace skb

using namespace skb;

extern "C" {

const char* skb_error_string(int code) {
    switch (code) {
        case SKB_OK: return "ok";
        case SKB_ERR_BAD_SHAPE: return "bad shape (sizes must be positive, M,N >= 2, BATCH/SYM need A == B, SYM needs M == N)";
        case SKB_ERR_BAD_ENUM: return "unknown enum value (static kind / scheme / pairs / arith / dtype)";
        case SKB_ERR_WORKSPACE: return "workspace missing or too small (see skb_*_workspace_bytes)";
        case SKB_ERR_UNSUPPORTED: return "shape not instantiated in this build";
        case SKB_ERR_CUDA: return "CUDA error (see skb_last_cuda_error)";
        case SKB_ERR_NULL: return "required pointer is NULL";
        default: return "unknown error code";
    }
}

int skb_last_cuda_error(void) { return g_last_cuda; }
int skb_version(void) { return 1; }
void skb_set_warps_per_sm(int warps) { set_warps_per_sm(warps); }

size_t skb_fwd_workspace_bytes(int A, int B, int M, int N, int D) {
    if (A <= 0 || B <= 0 || M <= 0 || N <= 0 || D <= 0) return 0;
    const size_t Dp = (size_t)padded_dim(D);
    return align256((size_t)A * M * Dp * sizeof(double)) + align256((size_t)B * N * Dp * sizeof(double));
}

int skb_sigkernel_fwd(const void* X, const void* Y, int io_dtype, int A, int B, int M, int N, int D,
                      int dyadic_order, int static_kind, double static_param, int scheme, int pairs,
                      int arith, double* out, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(A, B, M, N, dyadic_order, scheme, pairs, arith);
    if (rc) return rc;
    if (D <= 0) return SKB_ERR_BAD_SHAPE;
    if (static_kind != SKB_STATIC_LINEAR && static_kind != SKB_STATIC_RBF) return SKB_ERR_BAD_ENUM;
    if (io_dtype != SKB_F64 && io_dtype != SKB_F32) return SKB_ERR_BAD_ENUM;
    if (!X || !Y || !out) return SKB_ERR_NULL;
    if (!workspace || workspace_bytes < skb_fwd_workspace_bytes(A, B, M, N, D)) return SKB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int Dp = padded_dim(D);
    double* Xp = (double*)workspace;
    double* Yp = (double*)((char*)workspace + align256((size_t)A * M * Dp * sizeof(double)));
    double cx, nsc;
    prep_factors(static_kind, static_param, cx, nsc);
    rc = launch_prep(X, io_dtype, Xp, (long)A * M, D, Dp, cx, nsc, st);
    if (rc) return rc;
    rc = launch_prep(Y, io_dtype, Yp, (long)B * N, D, Dp, 1.0, nsc, st);
    if (rc) return rc;

    FwdArgs a;
    memset(&a, 0, sizeof(a));
    a.Xp = Xp; a.Yp = Yp; a.Ks = nullptr; a.out = out;
    a.njobs = njobs_of(A, B, pairs);
    a.A = A; a.B = B; a.M = M; a.N = N; a.Mv = M; a.Nv = N; a.Dp = Dp;
    a.kind = static_kind == SKB_STATIC_RBF ? 1 : 0;
    a.pairs = pairs;
    a.s1 = scheme == SKB_SCHEME_S1;
    a.scale4 = 1.0 / (double)(1ull << (2 * dyadic_order));
    return launch_forward(a, dyadic_order, arith == SKB_ARITH_EXACT, st);
}

int skb_sigkernel_fwd_from_static(const double* Ks, int A, int B, int M, int N, int dyadic_order, int scheme,
                                  int pairs, int arith, double* out, void* stream) {
    int rc = check_common(A, B, M, N, dyadic_order, scheme, pairs, arith);
    if (rc) return rc;
    if (!Ks || !out) return SKB_ERR_NULL;
    FwdArgs a;
    memset(&a, 0, sizeof(a));
    a.Ks = Ks; a.out = out;
    a.njobs = njobs_of(A, B, pairs);
    a.A = A; a.B = B; a.M = M; a.N = N; a.Mv = M; a.Nv = N; a.Dp = 0;
    a.kind = 2;
    a.pairs = pairs;
    a.s1 = scheme == SKB_SCHEME_S1;
    a.scale4 = 1.0 / (double)(1ull << (2 * dyadic_order));
    return launch_forward(a, dyadic_order, arith == SKB_ARITH_EXACT, (cudaStream_t)stream);
}

int skb_sigkernel_solve_increments(const double* inc, long P, int MM, int NN, int scheme, int arith,
                                   double* out, void* stream) {
    if (P <= 0 || MM < 1 || NN < 1 || P > 0x7fffffffL) return SKB_ERR_BAD_SHAPE;
    if (scheme != SKB_SCHEME_S2 && scheme != SKB_SCHEME_S1) return SKB_ERR_BAD_ENUM;
    if (arith != SKB_ARITH_FMA && arith != SKB_ARITH_EXACT) return SKB_ERR_BAD_ENUM;
    if (!inc || !out) return SKB_ERR_NULL;
    FwdArgs a;
    memset(&a, 0, sizeof(a));
    a.Ks = inc; a.out = out;
    a.njobs = P;
    a.A = (int)P; a.B = (int)P; a.M = MM + 1; a.N = NN + 1; a.Mv = MM; a.Nv = NN; a.Dp = 0;
    a.kind = 3;
    a.pairs = SKB_PAIRS_BATCH;
    a.s1 = scheme == SKB_SCHEME_S1;
    a.scale4 = 1.0;
    return launch_forward(a, 0, arith == SKB_ARITH_EXACT, (cudaStream_t)stream);
}

size_t skb_bwd_workspace_bytes(int A, int B, int M, int N, int D, int dyadic_order, int pairs) {
    (void)A; (void)B; (void)M; (void)N; (void)D; (void)dyadic_order; (void)pairs;
    return 0;
}

int skb_sigkernel_fwd_bwd(const void* X, const void* Y, int io_dtype, int A, int B, int M, int N, int D,
                          int dyadic_order, int static_kind, double static_param, int scheme, int pairs,
                          double* out, double* grad_points, void* workspace, size_t workspace_bytes,
                          void* stream) {
    (void)X; (void)Y; (void)io_dtype; (void)A; (void)B; (void)M; (void)N; (void)D; (void)dyadic_order;
    (void)static_kind; (void)static_param; (void)scheme; (void)pairs; (void)out; (void)grad_points;
    (void)workspace; (void)workspace_bytes; (void)stream;
    return SKB_ERR_UNSUPPORTED;
}

int skb_sigkernel_sensitivity_from_static(const double* Ks, int A, int B, int M, int N, int dyadic_order,
                                          int scheme, int pairs, double* out, double* S, void* workspace,
                                          size_t workspace_bytes, void* stream) {
    (void)Ks; (void)A; (void)B; (void)M; (void)N; (void)dyadic_order; (void)scheme; (void)pairs; (void)out;
    (void)S; (void)workspace; (void)workspace_bytes; (void)stream;
    return SKB_ERR_UNSUPPORTED;
}

}  // extern "C"
