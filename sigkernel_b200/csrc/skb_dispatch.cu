// skb_dispatch.cu -- path preparation kernel and the host-side dispatcher of solver_kernel.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "skb_common.cuh"
#include "skb_host.h"
#include "skb_tile.cuh"

namespace skb {

// X (batch, len, D) of type T -> Xp (batch, len, Dp) rows (nscale*|x|^2, c*x_0 .. c*x_{D-1}, 0 ...),
// optionally also the copy reversed along the length axis (the adjoint sweep solves the PDE of the
// reversed paths, sigkernel.py:434-438).
template <typename T>
__global__ void prep_kernel(const T* __restrict__ X, double* __restrict__ Xp, double* __restrict__ Xp_rev,
                            long batch, int len, int D, int Dp, double c, double nscale) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= batch * len) return;
    const T* x = X + r * D;
    const long bi = r / len;
    const int li = (int)(r - bi * len);
    double* o = Xp + r * Dp;
    double* orv = Xp_rev ? Xp_rev + (bi * len + (len - 1 - li)) * Dp : nullptr;
    double n = 0.0;
    for (int k = 0; k < D; ++k) {
        const double v = (double)x[k];
        n = fma(v, v, n);
        o[1 + k] = v * c;
        if (orv) orv[1 + k] = v * c;
    }
    o[0] = n * nscale;
    if (orv) orv[0] = n * nscale;
    for (int k = D + 1; k < Dp; ++k) {
        o[k] = 0.0;
        if (orv) orv[k] = 0.0;
    }
}

// X and Y in ONE launch, which also zeroes the job-queue counter of the solver launch that follows (a forward
// call is then two launches, not four: at 64x64 pairs of 32 points the launches were a third of the time)
template <typename T>
__global__ void prep2_kernel(const T* __restrict__ X, const T* __restrict__ Y, double* __restrict__ Xp,
                             double* __restrict__ Xr, double* __restrict__ Yp, double* __restrict__ Yr, long rowsX,
                             long rowsY, int M, int N, int D, int Dp, double cx, double nscale, unsigned int* counter) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0 && counter) { counter[0] = 0u; counter[32] = 0u; }   // job queue; finished-block count of the in-kernel rank barrier
    if (r >= rowsX + rowsY) return;
    const bool isx = r < rowsX;
    const long rr = isx ? r : r - rowsX;
    const int len = isx ? M : N;
    const T* x = (isx ? X : Y) + rr * D;
    double* P = isx ? Xp : Yp;
    double* Prv = isx ? Xr : Yr;
    const double c = isx ? cx : 1.0;
    const long bi = rr / len;
    const int li = (int)(rr - bi * len);
    double* o = P + rr * Dp;
    double* orv = Prv ? Prv + (bi * len + (len - 1 - li)) * Dp : nullptr;
    double n = 0.0;
    for (int k = 0; k < D; ++k) {
        const double v = (double)x[k];
        n = fma(v, v, n);
        o[1 + k] = v * c;
        if (orv) orv[1 + k] = v * c;
    }
    o[0] = n * nscale;
    if (orv) orv[0] = n * nscale;
    for (int k = D + 1; k < Dp; ++k) {
        o[k] = 0.0;
        if (orv) orv[k] = 0.0;
    }
}

int launch_prep2(const void* X, const void* Y, int dtype, double* Xp, double* Xr, double* Yp, double* Yr, long A, int M,
                 long B, int N, int D, int Dp, double cx, double nscale, unsigned int* counter, cudaStream_t st) {
    const long rowsX = A * M, rowsY = B * N;
    const int tb = 128;
    const unsigned grid = (unsigned)((rowsX + rowsY + tb - 1) / tb);
    if (dtype == SKB_F64)
        prep2_kernel<double><<<grid, tb, 0, st>>>((const double*)X, (const double*)Y, Xp, Xr, Yp, Yr, rowsX, rowsY, M, N, D, Dp, cx, nscale, counter);
    else
        prep2_kernel<float><<<grid, tb, 0, st>>>((const float*)X, (const float*)Y, Xp, Xr, Yp, Yr, rowsX, rowsY, M, N, D, Dp, cx, nscale, counter);
    return check_launch();
}

int launch_prep(const void* X, int dtype, double* Xp, double* Xp_rev, long batch, int len, int D, int Dp,
                double c, double nscale, cudaStream_t st) {
    const long rows = batch * len;
    if (rows == 0) return SKB_OK;
    const int tb = 128;
    const unsigned grid = (unsigned)((rows + tb - 1) / tb);
    if (dtype == SKB_F64)
        prep_kernel<double><<<grid, tb, 0, st>>>((const double*)X, Xp, Xp_rev, batch, len, D, Dp, c, nscale);
    else
        prep_kernel<float><<<grid, tb, 0, st>>>((const float*)X, Xp, Xp_rev, batch, len, D, Dp, c, nscale);
    return check_launch();
}

// ---- generic fallback helpers ---------------------------------------------------------------------
// static matrix of the fused kinds for jobs [job0, job0 + njobs): Ks[local job][i][j] = k(x_i, y_j)
// (libdevice exp here: this path is for shapes the fast kernels do not cover, not for speed)
__global__ void static_matrix_kernel(const double* __restrict__ Xp, const double* __restrict__ Yp,
                                     double* __restrict__ Ks, long job0, long njobs, int B, int M, int N,
                                     int Dp, int rbf, int batch) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per = (long)M * N;
    if (idx >= njobs * per) return;
    const long jl = idx / per;
    const int ij = (int)(idx - jl * per);
    const int i = ij / N, j = ij - i * N;
    const long pi = job0 + jl;
    const long a = batch ? pi : pi / B, b = batch ? pi : pi - a * B;
    const double* x = Xp + (a * M + i) * Dp;
    const double* y = Yp + (b * N + j) * Dp;
    double acc = x[0] + y[0];
    for (int k = 1; k < Dp; ++k) acc = fma(x[k], y[k], acc);
    Ks[idx] = rbf ? exp(acc) : acc;
}

// incc[job][i][j] = (((K[i+1,j+1] + K[i,j]) - K[i+1,j]) - K[i,j+1]) * 4^-d   (sigkernel.py:363-364)
__global__ void coarse_inc_kernel(const double* __restrict__ Ks, double* __restrict__ incc, long pairs, int M,
                                  int N, double scale4) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Mc = M - 1, Nc = N - 1;
    const long per = (long)Mc * Nc;
    if (idx >= pairs * per) return;
    const long p = idx / per;
    const int ij = (int)(idx - p * per);
    const int i = ij / Nc, j = ij - i * Nc;
    const double* K = Ks + p * ((long)M * N);
    const double k00 = K[(long)i * N + j], k01 = K[(long)i * N + j + 1];
    const double k10 = K[(long)(i + 1) * N + j], k11 = K[(long)(i + 1) * N + j + 1];
    incc[idx] = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(k11, k00), -k10), -k01), scale4);
}

int launch_static_matrix(const KArgs& a, int kind, long job0, long njobs, double* Ks, cudaStream_t st) {
    const long n = njobs * (long)a.M * a.N;
    if (n == 0) return SKB_OK;
    static_matrix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.Xp, a.Yp, Ks, job0, njobs, a.B, a.M, a.N, a.Dp,
                                                                      kind == KIND_RBF, a.pairs == PAIRS_BATCH);
    return check_launch();
}

int launch_coarse_increments(const double* Ks, double* incc, long pairs, int M, int N, double scale4, bool exact,
                             cudaStream_t st) {
    (void)exact;   // the kernel always uses the reference's rounding sequence
    const long n = pairs * (long)(M - 1) * (N - 1);
    if (n == 0) return SKB_OK;
    coarse_inc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Ks, incc, pairs, M, N, scale4);
    return check_launch();
}

static thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;
void set_profile_events(void* a, void* b) { g_ev_start = (cudaEvent_t)a; g_ev_stop = (cudaEvent_t)b; }

static int g_warps_per_sm = 0;
void set_warps_per_sm(int w) { g_warps_per_sm = w; }
int get_warps_per_sm() { return g_warps_per_sm; }

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// Row width of prepped paths: 1 norm slot + D coordinates, rounded up to one of the widths the
// fused kernels are specialised for (4, 6, 10 doubles), else to the next even number (generic loop).
int padded_dim(int D) {
    if (D + 1 <= 4) return 4;
    if (D + 1 <= 6) return 6;
    if (D + 1 <= 10) return 10;
    return (D + 2) & ~1;
}

static double scale4_of_logd(int d) { return 1.0 / (double)(1ull << (2 * d)); }

static int coarse_rows_per_lane(int M) {
    int rc = (M + 31) / 32, rcp = 1;
    while (rcp < rc) rcp <<= 1;
    return rcp;
}

int solver_rows_per_lane(int M, int logd) {
    const int rcp = coarse_rows_per_lane(M);
    if (rcp > 8 || logd > 5 || (rcp << logd) > 32) return -1;
    return rcp << logd;
}


// ---- fwd5: the forward-only kernel of the fused kinds (skb_fwd5.cuh) ---------------------------------
double fwd5_kscale(int logd) { return scale4_of_logd(logd) / sqrt(12.0); }

static bool fwd5_shape_ok(int rc, int logd) {   // SKB_FWD5_SHAPES of skb_fwd5.cuh
    return (rc == 1 && logd <= 3) || (rc == 2 && logd <= 2) || (rc == 4 && logd == 0);
}

static bool fwd5_r16_shape_ok(int rc, int logd) {   // SKB_FWD5_R16_SHAPES: one warp per pair only
    return (rc == 2 && logd == 3) || (rc == 4 && logd == 2) || (rc == 8 && logd == 1);
}

static bool fwd5_l16_shape_ok(int rc, int logd) {   // SKB_FWD5_L16_SHAPES of skb_fwd5.cuh
    return (rc == 1 && logd <= 3) || (rc == 2 && logd <= 3) || (rc == 4 && logd <= 2);
}

// warps per pair fwd5 would use (1, 2 or 4: the smallest count whose per-lane strip is an instantiated
// shape), the coarse rows per lane and the lanes per pair that go with it; 0 if the shape is not covered.
// Short paths (len_x <= 64) take the 16-lanes-per-pair variant when its strip is instantiated: twice the
// cells per lane and step for the same per-step overhead (measured 0.41 -> 0.35 ms at the headline config).
static int fwd5_plan(int M, int logd, int* rc_out, int* lpp_out = nullptr) {
    if (lpp_out) *lpp_out = 32;
    {
        int rc = (M + 15) / 16, rcp = 1;
        while (rcp < rc) rcp <<= 1;
        if (fwd5_l16_shape_ok(rcp, logd)) {
            if (rc_out) *rc_out = rcp;
            if (lpp_out) *lpp_out = 16;
            return 1;
        }
    }
    for (int nw = 1; nw <= 4; nw *= 2) {
        int rc = (M + 32 * nw - 1) / (32 * nw), rcp = 1;
        while (rcp < rc) rcp <<= 1;
        if (fwd5_shape_ok(rcp, logd) || (nw == 1 && fwd5_r16_shape_ok(rcp, logd))) {
            if (rc_out) *rc_out = rcp;
            return nw;
        }
    }
    return 0;
}

int fwd5_warps_per_pair(int M, int logd, int* lpp) { return fwd5_plan(M, logd, nullptr, lpp); }

// true if the fwd5 variant of this shape evaluates exp with the pre-scaled argument (SCALED in skb_fwd5.cuh)
bool fwd5_scaled_exp(int M, int logd, int D) {
    int rcp = 0, lpp = 32;
    const int nw = fwd5_plan(M, logd, &rcp, &lpp);
    if (nw != 1) return false;
    const int dp2 = padded_dim(D) / 2, R = rcp << logd;
    return R > 8 && rcp * dp2 <= ((lpp == 16 && R > 8) ? 12 : 8);
}

bool fwd5_applies(int kind, int M, int N, int D, int logd, bool s1) {
    if (N < 4) return false;
    if (kind != KIND_RBF && kind != KIND_LINEAR) return false;
    const int Dp = padded_dim(D);
    if (Dp != 4 && Dp != 6 && Dp != 10) return false;
    const int nw = fwd5_plan(M, logd, nullptr);
    return s1 ? nw == 1 : nw > 0;           // scheme S1 is instantiated for the single-warp variants
}

// 2^(j/2048), j = 0..2047, in device memory: every block of the RBF kernels copies (a stride of) it, times kscale, into
// its shared exp table instead of evaluating exp2 64 times per lane.  Built once per device on first use (host libm,
// synchronous copy); the pointer stays valid for the life of the process.
static const double* device_exp_table() {
    static std::mutex mu;
    static const double* tabs[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (tabs[dev]) return tabs[dev];
    static double host[2048];
    for (int j = 0; j < 2048; ++j) host[j] = exp2((double)j / 2048.0);
    double* d = nullptr;
    if (cudaMalloc(&d, sizeof(host)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaMemcpy(d, host, sizeof(host), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); cudaFree(d); return nullptr; }
    tabs[dev] = d;
    return d;
}

static int fill_v5_constants(KArgs& args, int logd) {
    args.exp_tab = device_exp_table();
    if (!args.exp_tab) return SKB_ERR_CUDA;
    args.kscale = fwd5_kscale(logd);
    args.inv_kscale = 1.0 / args.kscale;
    args.sqrt3 = sqrt(3.0);
    args.ek = 369.32993046757463;                 // 256 / ln 2
    args.ehi = -0x1.62e42fee00000p-9;             // ln2/256 = hi + lo; hi has 21 trailing zero bits
    args.elo = -0x1.a39ef35793c76p-41;
    args.e4 = 1.0 / 24.0;
    args.e3 = 1.0 / 6.0;
    return SKB_OK;
}

// SKB_ADJ5_SHAPES of skb_fwd5.cuh: one warp per pair, dyadic order >= 1, <= 8 fine rows per lane
static bool adjoint5_shape_ok(int rc, int logd) { return (rc == 1 && logd >= 1 && logd <= 3) || (rc == 2 && logd >= 1 && logd <= 2); }

bool adjoint5_applies(int kind, int M, int N, int D, int logd, bool s1) {
    if (s1 || N < 4) return false;
    if (kind != KIND_RBF && kind != KIND_LINEAR) return false;
    const int Dp = padded_dim(D);
    if (Dp != 4 && Dp != 6 && Dp != 10) return false;
    if (solver_rows_per_lane(M, logd) < 0) return false;
    return adjoint5_shape_ok(coarse_rows_per_lane(M), logd);
}

int launch_adjoint5(int mode, int kind, int logd, KArgs args, cudaStream_t st) {
    const int rcp = coarse_rows_per_lane(args.M);
    if (int e = fill_v5_constants(args, logd)) return e;
    args.pitch = 32L * (rcp << logd);
    if (!args.counter) return SKB_ERR_WORKSPACE;
    int rc = args.counter_clean ? SKB_OK : check_cuda(cudaMemsetAsync(args.counter, 0, sizeof(unsigned int), st));
    if (rc) return rc;
    const bool rbf = kind == KIND_RBF;
    if (mode == 1) return rbf ? launch_group_adj5_rbf_store(rcp, logd, args.Dp / 2, args, st) : launch_group_adj5_lin_store(rcp, logd, args.Dp / 2, args, st);
    if (mode == 3) return rbf ? launch_group_adj5_rbf_rev(rcp, logd, args.Dp / 2, args, st) : launch_group_adj5_lin_rev(rcp, logd, args.Dp / 2, args, st);
    return SKB_ERR_UNSUPPORTED;
}

// ---- adjoint by reconstruction ---------------------------------------------------------------------------------
static int g_adjoint_mode = -1;
void set_adjoint_mode(int mode) { g_adjoint_mode = mode; }
int get_adjoint_mode() { return g_adjoint_mode; }

static int pow2_ceil(int v) {
    int r = 1;
    while (r < v) r <<= 1;
    return r;
}

// warps per pair (1, 2, 4; 0 = not covered), coarse rows per lane and lanes per pair of the reconstruction kernels:
// SKB_RECON5_*_SHAPES of skb_recon5_launch.cuh (strips of at most 8 fine rows)
static int recon5_plan(int mode, int M, int logd, int* rc_out, int* lpp_out) {
    if (logd > 3) return 0;
    const int rmax = 8 >> logd;
    if (mode == MODE_REV_RECON_SYM) {
        // the unordered-pair sweep: one warp of 32 lanes per pair only
        const int rc = pow2_ceil((M + 31) / 32);
        *lpp_out = 32;
        *rc_out = rc;
        return rc <= rmax ? 1 : 0;
    }
    // 16 lanes per pair (two pair streams per warp) wherever the strip is instantiated: the forward pass always, the
    // reversed sweep only on request (mode 2; measured at 128 x 128 pairs of 64 points, dyadic order 1: 0.71 ms with 32
    // lanes per pair, 0.89 ms with 16 -- twice the rows per lane do not pay for the larger, slower step there)
    if (mode == MODE_FWD_EMIT || g_adjoint_mode == 2) {
        const int rc = pow2_ceil((M + 15) / 16);
        if (rc <= rmax && (rc << logd) >= 4) {
            *rc_out = rc; *lpp_out = 16;
            return 1;
        }
    }
    *lpp_out = 32;
    {
        const int rc = pow2_ceil((M + 31) / 32);
        if (rc <= rmax) {
            *rc_out = rc;
            return 1;
        }
    }
    if (logd <= 2) {
        for (int nw = 2; nw <= 4; nw *= 2)
            if (M <= 32 * nw * rmax) {
                *rc_out = rmax;
                return nw;
            }
    }
    return 0;
}

// the unordered-pair sweep of Gram(X, X) (MODE_REV_RECON_SYM): single-warp strips with register accumulators
bool recon5_sym_applies(int kind, int M, int N, int D, int logd, bool s1) {
    if (g_adjoint_mode == 0 || g_adjoint_mode == 3 || s1 || N < 4 || M != N) return false;
    if (kind != KIND_RBF && kind != KIND_LINEAR) return false;
    const int Dp = padded_dim(D);
    if (Dp != 4 && Dp != 6 && Dp != 10) return false;
    int rc, lpp;
    if (recon5_plan(MODE_REV_RECON_SYM, M, logd, &rc, &lpp) != 1) return false;
    return rc * (Dp / 2) <= 6;
}

bool recon5_applies(int kind, int M, int N, int D, int logd, bool s1) {
    if (g_adjoint_mode == 0 || s1 || N < 4) return false;
    if (kind != KIND_RBF && kind != KIND_LINEAR) return false;
    const int Dp = padded_dim(D);
    if (Dp != 4 && Dp != 6 && Dp != 10) return false;
    int rc, lpp;
    return recon5_plan(MODE_REV_RECON, M, logd, &rc, &lpp) > 0;
}

int launch_recon5(int mode, int kind, int logd, KArgs args, cudaStream_t st) {
    int rc = 0, lpp = 32;
    const int nw = recon5_plan(mode, args.M, logd, &rc, &lpp);
    if (nw == 0) return SKB_ERR_UNSUPPORTED;
    if (int e = fill_v5_constants(args, logd)) return e;
    {
        // parked-sum buffers of the reversed sweep: the first lane of a warp is up to 31 steps (+ the step of the boundary
        // check) ahead of the flush, i.e. ceil(34 / N) pairs
        int nbuf = 1;
        while (nbuf * args.N < 34) nbuf *= 2;
        args.fbuf_mask = nbuf - 1;
    }
    if (!args.counter) return SKB_ERR_WORKSPACE;
    int err = args.counter_clean ? SKB_OK : check_cuda(cudaMemsetAsync(args.counter, 0, sizeof(unsigned int), st));
    if (err) return err;
    const bool rbf = kind == KIND_RBF;
    if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_start, st);
    if (nw > 1) err = rbf ? launch_group_recon5_rbf_nw(mode, rc, logd, args.Dp / 2, nw, args, st) : launch_group_recon5_lin_nw(mode, rc, logd, args.Dp / 2, nw, args, st);
    else if (lpp == 16) err = rbf ? launch_group_recon5_rbf_l16(mode, rc, logd, args.Dp / 2, 1, args, st) : launch_group_recon5_lin_l16(mode, rc, logd, args.Dp / 2, 1, args, st);
    else err = rbf ? launch_group_recon5_rbf_l32(mode, rc, logd, args.Dp / 2, 1, args, st) : launch_group_recon5_lin_l32(mode, rc, logd, args.Dp / 2, 1, args, st);
    if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_stop, st);
    return err;
}

__global__ void vjp_accumulate_kernel(const double* __restrict__ gp, long job0, long njobs, int A, int B, int M, int D, int pairs,
                                      const double* __restrict__ gout, double w_diag, double w_off, double* __restrict__ gradX,
                                      const unsigned int* cond) {
    if (cond != nullptr && *cond == 0u) return;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per = (long)M * D;
    if (idx >= njobs * per) return;
    const long jl = idx / per;
    const long md = idx - jl * per;
    const long pi = job0 + jl;
    const long a = pairs == PAIRS_BATCH ? pi : pi / B;
    const long b = pairs == PAIRS_BATCH ? pi : pi - a * B;
    const double coef = gout ? gout[pi] : (a == b ? w_diag : w_off);
    (void)A;
    atomicAdd(gradX + a * per + md, coef * gp[idx]);
}

int launch_vjp_accumulate(const double* gp, long job0, long njobs, int A, int B, int M, int D, int pairs, const double* gout,
                          double w_diag, double w_off, double* gradX, const unsigned int* cond, cudaStream_t st) {
    const long n = njobs * (long)M * D;
    if (n == 0) return SKB_OK;
    vjp_accumulate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gp, job0, njobs, A, B, M, D, pairs, gout, w_diag, w_off, gradX, cond);
    return check_launch();
}

__global__ void cond_zero_kernel(double* __restrict__ ptr, size_t n, const unsigned int* cond) {
    if (cond != nullptr && *cond == 0u) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) ptr[i] = 0.0;
}

int launch_cond_zero(double* ptr, size_t n, const unsigned int* cond, cudaStream_t st) {
    if (n == 0) return SKB_OK;
    size_t blocks = (n + 255) / 256;
    if (blocks > 2048) blocks = 2048;
    cond_zero_kernel<<<(unsigned)blocks, 256, 0, st>>>(ptr, n, cond);
    return check_launch();
}

__global__ void rank_barrier_kernel(const KArgs p) { rank_barrier(p); }

int launch_rank_barrier(const KArgs& a, cudaStream_t st) {
    rank_barrier_kernel<<<1, 32, 0, st>>>(a);
    return check_launch();
}

int launch_forward5(int kind, int logd, KArgs args, cudaStream_t st) {
    int rcp = 0, lpp = 32;
    const int nw = fwd5_plan(args.M, logd, &rcp, &lpp);
    if (nw == 0) return SKB_ERR_UNSUPPORTED;
    if (int e = fill_v5_constants(args, logd)) return e;
    if (fwd5_scaled_exp(args.M, logd, args.D)) {
        // single-warp forward variants with the x rows in registers: exp argument in units of c = ln2 / 2048 (exp_scaled5): e^(r c) - 1 = r (c + r (c^2/2 + r c^3/6))
        const double c = 0.693147180559945309417232121458 / 2048.0;
        args.ek = c; args.e4 = c * c / 2.0; args.e3 = c * c * c / 6.0;
    }
    if (!args.counter) return SKB_ERR_WORKSPACE;
    int rc = args.counter_clean ? SKB_OK : check_cuda(cudaMemsetAsync(args.counter, 0, sizeof(unsigned int), st));
    if (rc) return rc;
    if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_start, st);
    typedef int (*fwd5_fn)(int, int, int, const KArgs&, cudaStream_t);
    static const fwd5_fn table[2][3] = {
        {launch_group_fwd5_lin_nw1, launch_group_fwd5_lin_nw2, launch_group_fwd5_lin_nw4},
        {launch_group_fwd5_rbf_nw1, launch_group_fwd5_rbf_nw2, launch_group_fwd5_rbf_nw4}};
    if (lpp == 16)
        rc = kind == KIND_RBF ? launch_group_fwd5_rbf_l16(rcp, logd, args.Dp / 2, args, st)
                              : launch_group_fwd5_lin_l16(rcp, logd, args.Dp / 2, args, st);
    else
        rc = table[kind == KIND_RBF ? 1 : 0][nw == 1 ? 0 : (nw == 2 ? 1 : 2)](rcp, logd, args.Dp / 2, args, st);
    if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_stop, st);
    return rc;
}

// ---- tile forward kernel (skb_tile.cuh) --------------------------------------------------------------------
static int g_tile_mode = -1;
void set_tile_mode(int mode) { g_tile_mode = mode; }

static const int kTileW = 8, kTileR = 16;     // warps per block (strips per band), fine rows per strip

static bool tile_shape_ok(int logd) { return logd >= 1 && logd <= 3; }   // SKB_TILE_SHAPES: RC = 16 >> logd

static long tile_count(int A, int B, int pairs) {
    const long nta = (A + 31) / 32;
    return pairs == PAIRS_BATCH ? nta : nta * (long)B;
}

bool tile_applies(int kind, int A, int B, int M, int N, int D, int logd, bool s1, int pairs) {
    if (g_tile_mode == 0 || s1 || N < 16) return false;     // N - 1 >= TILE_RD_MAX + 1 (skb_tile.cuh)
    if (kind != KIND_RBF && kind != KIND_LINEAR) return false;
    if (pairs != PAIRS_GRAM && pairs != PAIRS_BATCH) return false;
    const int Dp = padded_dim(D);
    if (Dp != 4 && Dp != 6 && Dp != 10) return false;
    if (!tile_shape_ok(logd)) return false;
    const long MMf = (long)(M - 1) << logd;
    const long strips = (MMf + kTileR - 1) / kTileR;
    const long ntiles = tile_count(A, B, pairs);
    const long nbands = (strips + kTileW - 1) / kTileW;
    if (ntiles * nbands > 0x3fffffffL) return false;
    // opt-in only: at every BASELINE config the tile kernel is slower than fwd5_kernel so far
    return g_tile_mode == 1;
}

double tile_arg_scale() { return 2048.0 / 0.693147180559945309417232121458; }

static size_t tile_align(size_t x) { return (x + 255) & ~(size_t)255; }

size_t tile_workspace_bytes(int A, int B, int M, int N, int logd, int pairs) {
    const long MMf = (long)(M - 1) << logd;
    const long strips = (MMf + kTileR - 1) / kTileR;
    const size_t ntiles = (size_t)tile_count(A, B, pairs), nbands = (size_t)((strips + kTileW - 1) / kTileW);
    const size_t F = (size_t)1 << logd, NS = (size_t)(N - 1);
    return tile_align(ntiles * NS * 32 * sizeof(double)) + tile_align(ntiles * (nbands - 1) * NS * (F / 2 + 1) * 32 * sizeof(double2)) +
           tile_align(ntiles * (nbands - 1) * sizeof(unsigned int) + 4);
}

template <int KIND>
__global__ void tile_d0_zero_kernel(const TArgs p, int Dp, double inv_s, long nready);

int launch_tile_forward(int kind, int logd, const KArgs& a, void* tile_ws, cudaStream_t st) {
    if (!tile_shape_ok(logd) || !tile_ws || !a.counter) return SKB_ERR_UNSUPPORTED;
    TArgs t;
    memset(&t, 0, sizeof(t));
    const long MMf = (long)(a.M - 1) << logd;
    const long strips = (MMf + kTileR - 1) / kTileR;
    t.Xp = a.Xp; t.Yp = a.Yp; t.out = a.out; t.counter = a.counter;
    t.A = a.A; t.B = a.B; t.M = a.M; t.N = a.N; t.pairs = a.pairs;
    t.nta = (a.A + 31) / 32;
    t.ntiles = (int)tile_count(a.A, a.B, a.pairs);
    t.nbands = (int)((strips + kTileW - 1) / kTileW);
    t.njobs = t.ntiles * t.nbands;
    t.kscale = fwd5_kscale(logd);
    t.sqrt3 = sqrt(3.0);
    const double c = 0.693147180559945309417232121458 / 2048.0;
    t.c1 = c; t.c2 = c * c / 2.0; t.c3 = c * c * c / 6.0;
    const size_t F = (size_t)1 << logd, NS = (size_t)(a.N - 1);
    char* w = (char*)tile_ws;
    t.d0 = (const double*)w;
    w += tile_align((size_t)t.ntiles * NS * 32 * sizeof(double));
    t.bnd = (double2*)w;
    w += tile_align((size_t)t.ntiles * (t.nbands - 1) * NS * (F / 2 + 1) * 32 * sizeof(double2));
    t.ready = (unsigned int*)w;
    int rc = a.counter_clean ? SKB_OK : check_cuda(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), st));
    if (rc) return rc;
    {
        // top boundary of band 0 (+ zeroing of the band hand-off counters)
        const long n = (long)t.ntiles * (long)NS * 32;
        const long nready = (long)t.ntiles * (t.nbands - 1);
        const unsigned grid = (unsigned)(((n > nready ? n : nready) + 255) / 256);
        if (kind == KIND_RBF) tile_d0_zero_kernel<KIND_RBF><<<grid, 256, 0, st>>>(t, a.Dp, 1.0 / tile_arg_scale(), nready);
        else tile_d0_zero_kernel<KIND_LINEAR><<<grid, 256, 0, st>>>(t, a.Dp, 1.0, nready);
        rc = check_launch();
        if (rc) return rc;
    }
    if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_start, st);
    rc = kind == KIND_RBF ? launch_group_tile_rbf(kTileR >> logd, logd, a.Dp / 2, t, st)
                          : launch_group_tile_lin(kTileR >> logd, logd, a.Dp / 2, t, st);
    if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_stop, st);
    return rc;
}

template <int KIND>
__global__ void tile_d0_zero_kernel(const TArgs p, int Dp, double inv_s, long nready) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < nready) p.ready[idx] = 0u;
    const int lane = (int)(idx & 31);
    const long tc = idx >> 5;
    const int NS = p.N - 1;
    if (tc >= (long)p.ntiles * NS) return;
    const int tile = (int)(tc / NS), c = (int)(tc - (long)tile * NS);
    int a, b;
    if (p.pairs == PAIRS_BATCH) {
        a = b = tile * 32 + lane;
    } else {
        b = tile / p.nta;
        a = (tile - b * p.nta) * 32 + lane;
    }
    a = a < p.A ? a : p.A - 1;
    b = b < p.B ? b : p.B - 1;
    const double* x = p.Xp + ((size_t)a * p.M) * Dp;
    const double* y0 = p.Yp + ((size_t)b * p.N + c) * Dp;
    const double* y1 = y0 + Dp;
    double a0 = x[0] + y0[0], a1 = x[0] + y1[0];
    for (int k = 1; k < Dp; ++k) {
        a0 = fma(x[k], y0[k], a0);
        a1 = fma(x[k], y1[k], a1);
    }
    if (KIND == KIND_RBF) {
        a0 = p.kscale * exp(a0 * inv_s);
        a1 = p.kscale * exp(a1 * inv_s);
    }
    const_cast<double*>(p.d0)[idx] = a1 - a0;
}

int launch_solver(int mode, int kind, int logd, bool exact, KArgs args, cudaStream_t st) {
    if (kind == KIND_INCV) logd = 0;   // the band sweep runs on the fine grid
    if (solver_rows_per_lane(args.M, logd) < 0) return SKB_ERR_UNSUPPORTED;
    const int rcp = coarse_rows_per_lane(args.M);
    args.tstar = (args.M - 2) / rcp;
    args.rcstar = (args.M - 2) % rcp;
    args.pitch = 32L * (rcp << logd);
    if (!args.counter) return SKB_ERR_WORKSPACE;
    int rc = args.counter_clean ? SKB_OK : check_cuda(cudaMemsetAsync(args.counter, 0, sizeof(unsigned int), st));
    if (rc) return rc;
    int dp2 = 0;
    if (kind == KIND_RBF || kind == KIND_LINEAR) {
        if (args.Dp == 4 || args.Dp == 6 || args.Dp == 10) dp2 = args.Dp / 2;
    }
    group_fn fn = nullptr;
    const bool rbf = kind == KIND_RBF, lin = kind == KIND_LINEAR;
    if (kind == KIND_STATIC || kind == KIND_INC || kind == KIND_INCV) fn = launch_group_static;
    else if (mode == 0) fn = rbf ? launch_group_fwd_rbf : (lin ? launch_group_fwd_lin : nullptr);
    else if (mode == 1) fn = rbf ? launch_group_store_rbf : (lin ? launch_group_store_lin : nullptr);
    else if (mode == 3) fn = rbf ? launch_group_rev_rbf : (lin ? launch_group_rev_lin : nullptr);
    if (!fn) return SKB_ERR_UNSUPPORTED;
    if (exact && (rbf || lin)) return SKB_ERR_UNSUPPORTED;
    if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_start, st);
    rc = fn(mode, kind, rcp, logd, dp2, exact, args, st);
    if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_stop, st);
    return rc;
}

}  // namespace skb
