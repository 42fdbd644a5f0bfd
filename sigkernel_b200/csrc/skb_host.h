// skb_host.h -- host-side declarations shared by the translation units of libsigkernel_b200.so
#pragma once
#include <cuda_runtime.h>
#include "../../include/sigkernel_b200.h"

namespace skb {

// Arguments of solver_kernel (skb_solver.cuh).  Plain data, passed by value.
struct KArgs {
    const double* Xp;      // prepped X rows [A*M][Dp]: (nx, c*x_0 .. c*x_{D-1}, 0 pad); reversed copy in REV modes
    const double* Yp;      // prepped Y rows [B*N][Dp]: (ny,   y_0 ..   y_{D-1}, 0 pad); reversed copy in REV modes
    const double* Ks;      // KIND_STATIC: coarse static matrix; KIND_INC: fine increments
    double* out;           // FWD modes: k(X_a, Y_b)
    double* scratch;       // FWD_STORE writes / REV reads the forward grid: [job][q][pitch], behind a front pad
    double* S;             // REV_S: coarse sensitivities (pairs, M-1, N-1)
    double* grad;          // REV_GRAD: per-point gradients (pairs, M, D)
    unsigned int* counter; // job queue (zeroed before the launch)
    int counter_clean;     // host side: the counter was already zeroed by the preparation kernel (first launch of a call)
    long job0;             // first job of this launch in the pair enumeration
    long pitch;            // scratch row pitch in doubles (32 * R)
    int njobs;             // jobs of this launch
    int A, B;
    int M, N;              // production rows / columns per pair (nodes; KIND_INC: MM+1, NN+1)
    int Mv, Nv;            // rows / columns stored in Ks
    int Dp, D;             // doubles per prepped row (even); path dimension
    int pairs, s1;
    int tstar, rcstar;     // lane / coarse row that ends up holding u[MM,NN]
    // KIND_INCV (generic row-band fallback): the launch sweeps fine rows [band_row0, band_row0 + M - 1)
    const double* band_top; // u of the fine row above the band, per pair and fine column (NULL: boundary 1)
    double* band_bot;       // receives u of the band's last fine row (NULL: last band)
    int band_row0;          // first fine row of the band
    int dshift;             // dyadic order (coarse index = fine index >> dshift)
    int Mc, Nc;             // coarse increment matrix dims
    double scale4;         // 4^-d (dyadic refinement: tile()/2^d twice, sigkernel.py:364)
    double gscale;         // REV_GRAD: 2/sigma (RBF) or the linear scale factor
    // fwd5 (skb_fwd5.cuh): the static kernel is produced pre-scaled by kscale = 4^-d / sqrt(12); constants of
    // the coefficient polynomial and of the table-driven exp live here so that they are constant-bank operands
    double kscale, sqrt3, inv_kscale;
    double ek, ehi, elo, e4, e3;
    const double* exp_tab;   // 2^(j/2048), j < 2048 (device_exp_table(), skb_dispatch.cu)
    // adjoint by reconstruction (MODE_FWD_EMIT / MODE_REV_RECON of skb_fwd5.cuh): the forward pass leaves the last row
    // and the last column of every pair's grid, brow[slot][k] = u[MM, k] (k = 0..NN) and bcol[slot][k] = u[k, NN]
    // (k = 0..MM); the reversed sweep rebuilds the grid backwards from them.  slot = job index of the forward launch.
    double* brow;
    double* bcol;
    long brow_stride, bcol_stride;   // doubles per pair (multiples of 4)
    int bsym;                        // REV_RECON: the boundaries come from a PAIRS_SYM forward (slot of (a,b) = triangle
                                     // index of (min, max); a > b reads the transposed grid: brow <-> bcol)
    unsigned int* flag;              // REV_RECON: set to 1 if a rebuilt grid misses u[., 0] = 1 by more than recon_tol
    double recon_tol;
    // in-kernel rank barrier of the sharded forward (skb_sigkernel_fwd_range): the last block to finish stores sig_epoch to
    // slot sig_rank of every rank's signal array (peer memory) and waits until every slot of its own array has reached it
    unsigned long long* sig_peer[8];
    int sig_rank;
    unsigned long long sig_epoch;    // 0: no barrier
    int fbuf_mask;                   // REV_RECON: parked-sum buffers - 1 (a power of two with N * buffers >= 34; skb_fwd5.cuh UFLUSH)
    const unsigned int* cond;        // if non-NULL the kernel returns at once unless *cond != 0 (stored-grid fallback)
    // fused loss head of REV_RECON: gradX[a, m, :] += coef(a,b) * grad_points[a, b, m, :] (atomic), with
    // coef = gout[pair] if gout, else (a == b ? w_diag : w_off)
    const double* gout;
    double* gradX;
    double w_diag, w_off;
    // multi-GPU gather without a collective (forward, GRAM / BATCH): when n_peer > 0 every result is stored to each of
    // out_peer[0 .. n_peer) (this rank's block inside every rank's copy of G, peer memory over NVLink) instead of `out`
    double* out_peer[8];
    int n_peer;
};

// records the cudaError_t for skb_last_cuda_error(); returns SKB_OK or SKB_ERR_CUDA
int check_cuda(cudaError_t e);
inline int check_launch() { return check_cuda(cudaGetLastError()); }

void set_warps_per_sm(int w);
void set_profile_events(void* start, void* stop);
int get_warps_per_sm();
int sm_count();

int launch_prep2(const void* X, const void* Y, int dtype, double* Xp, double* Xr, double* Yp, double* Yr, long A, int M,
                 long B, int N, int D, int Dp, double cx, double nscale, unsigned int* counter, cudaStream_t st);
int launch_prep(const void* X, int dtype, double* Xp, double* Xp_rev, long batch, int len, int D, int Dp,
                double c, double nscale, cudaStream_t st);

// mode: MODE_* of skb_solver.cuh; kind: KIND_*; logd: dyadic order; dp2: Dp/2 specialisation (0 = generic).
// Picks RC from args.M, fills tstar/rcstar, zeroes the job counter, launches.
int launch_solver(int mode, int kind, int logd, bool exact, KArgs args, cudaStream_t st);
// rows-per-lane the dispatcher would use (needed to size the scratch pitch); <0 if unsupported
int solver_rows_per_lane(int M, int logd);
// generic fallback: coarse increments (pairs, M-1, N-1) from a static matrix; static matrix of the fused kinds
int launch_coarse_increments(const double* Ks, double* incc, long pairs, int M, int N, double scale4, bool exact, cudaStream_t st);
int launch_static_matrix(const KArgs& a, int kind, long job0, long njobs, double* Ks, cudaStream_t st);
bool recon5_sym_applies(int kind, int M, int N, int D, int logd, bool s1);   // MODE_REV_RECON_SYM instantiated for the shape
void set_deriv_mode(int mode);                                   // skb_deriv.cu: 0 = bit-exact diagonal kernel always
int launch_rank_barrier(const KArgs& a, cudaStream_t st);      // the in-kernel rank barrier on its own (a rank without pairs)
// backward on materialised grids (skb_generic_adj.cu)
int launch_grid_solve(const double* inc, double* U, double* out, long job0, long njobs, int M, int N, int d, bool s1, cudaStream_t st);
int launch_coarse_sens(const double* U, double* S, long njobs, int M, int N, int d, double scale4, cudaStream_t st);
int launch_grad_from_sens(const double* S, const double* Ks, const KArgs& a, int kind, double* grad, long job0, long njobs, cudaStream_t st);
// padded row width the fused kinds are specialised for
int padded_dim(int D);

// kernel + directional derivatives (skb_deriv.cu): K0, K1, K2 coarse static matrices (pairs, M, N);
// inc3 scratch (pairs, M-1, N-1, 3); out3 (pairs, 3)
int launch_derivatives(const double* K0, const double* K1, const double* K2, long pairs, int M, int N, int d,
                       double eps, double* inc3, double* out3, cudaStream_t st);

// per-group launchers (one translation unit each, to parallelise compilation)
typedef int (*group_fn)(int mode, int kind, int rc, int logd, int dp2, bool exact, const KArgs&, cudaStream_t);
int launch_group_fwd_rbf(int, int, int, int, int, bool, const KArgs&, cudaStream_t);
int launch_group_fwd_lin(int, int, int, int, int, bool, const KArgs&, cudaStream_t);
int launch_group_static(int, int, int, int, int, bool, const KArgs&, cudaStream_t);
int launch_group_store_rbf(int, int, int, int, int, bool, const KArgs&, cudaStream_t);
int launch_group_store_lin(int, int, int, int, int, bool, const KArgs&, cudaStream_t);
int launch_group_rev_rbf(int, int, int, int, int, bool, const KArgs&, cudaStream_t);
int launch_group_rev_lin(int, int, int, int, int, bool, const KArgs&, cudaStream_t);
// forward-only kernel of the fused kinds (skb_fwd5.cuh); SKB_ERR_UNSUPPORTED if the shape is not instantiated
// (one translation unit per static kind and warps-per-pair count NW)
int launch_group_fwd5_rbf_nw1(int rc, int logd, int dp2, const KArgs&, cudaStream_t);
int launch_group_fwd5_rbf_nw2(int rc, int logd, int dp2, const KArgs&, cudaStream_t);
int launch_group_fwd5_rbf_nw4(int rc, int logd, int dp2, const KArgs&, cudaStream_t);
int launch_group_fwd5_lin_nw1(int rc, int logd, int dp2, const KArgs&, cudaStream_t);
int launch_group_fwd5_lin_nw2(int rc, int logd, int dp2, const KArgs&, cudaStream_t);
int launch_group_fwd5_lin_nw4(int rc, int logd, int dp2, const KArgs&, cudaStream_t);
int launch_group_fwd5_rbf_l16(int rc, int logd, int dp2, const KArgs&, cudaStream_t);
int launch_group_fwd5_lin_l16(int rc, int logd, int dp2, const KArgs&, cudaStream_t);
// true if skb_sigkernel_fwd should take the fwd5 path for this problem (scheme S2, N >= 4, shape instantiated)
bool fwd5_applies(int kind, int M, int N, int D, int logd, bool s1);
// warps per pair fwd5 would use (0: not covered); *lpp (if given) receives the lanes per pair (32 or 16)
int fwd5_warps_per_pair(int M, int logd, int* lpp = nullptr);
bool fwd5_scaled_exp(int M, int logd, int D);
// scale the static kernel is produced with on the fwd5 path (Linear: folded into the prepared X rows)
double fwd5_kscale(int logd);
int launch_forward5(int kind, int logd, KArgs args, cudaStream_t st);
// adjoint passes on the v5 kernel (MODE_FWD_STORE / MODE_REV_GRAD of skb_fwd5.cuh): one warp per pair, even strips
bool adjoint5_applies(int kind, int M, int N, int D, int logd, bool s1);
int launch_adjoint5(int mode, int kind, int logd, KArgs args, cudaStream_t st);
int launch_group_adj5_rbf_store(int rc, int logd, int dp2, const KArgs&, cudaStream_t);
int launch_group_adj5_rbf_rev(int rc, int logd, int dp2, const KArgs&, cudaStream_t);
int launch_group_adj5_lin_store(int rc, int logd, int dp2, const KArgs&, cudaStream_t);
int launch_group_adj5_lin_rev(int rc, int logd, int dp2, const KArgs&, cudaStream_t);

// ---- adjoint by reconstruction (MODE_FWD_EMIT / MODE_REV_RECON of skb_fwd5.cuh) ---------------------------------
// development / tuning knob (process-wide): -1 / 1 = default (reconstruction, 32 lanes per pair), 0 = stored-grid
// kernels only, 2 = reconstruction with 16 lanes per pair where instantiated, 3 = default without the unordered-pair sweep
void set_adjoint_mode(int mode);
int get_adjoint_mode();
// true if the reconstruction kernels cover the problem (fused kind, scheme S2, len_y >= 4, strips of <= 8 fine rows on
// up to 4 warps per pair: (len_x - 1) 2^d <= 1024 at dyadic order <= 2)
bool recon5_applies(int kind, int M, int N, int D, int logd, bool s1);
// mode: MODE_FWD_EMIT or MODE_REV_RECON; args as for launch_forward5 plus the boundary / loss-head fields
int launch_recon5(int mode, int kind, int logd, KArgs args, cudaStream_t st);
int launch_group_recon5_rbf_l16(int mode, int rc, int logd, int dp2, int nw, const KArgs&, cudaStream_t);
int launch_group_recon5_rbf_l32(int mode, int rc, int logd, int dp2, int nw, const KArgs&, cudaStream_t);
int launch_group_recon5_rbf_nw(int mode, int rc, int logd, int dp2, int nw, const KArgs&, cudaStream_t);
int launch_group_recon5_lin_l16(int mode, int rc, int logd, int dp2, int nw, const KArgs&, cudaStream_t);
int launch_group_recon5_lin_l32(int mode, int rc, int logd, int dp2, int nw, const KArgs&, cudaStream_t);
int launch_group_recon5_lin_nw(int mode, int rc, int logd, int dp2, int nw, const KArgs&, cudaStream_t);
// gradX[a, m, :] += coef(a, b) * gp[job - job0, m, :] for jobs [job0, job0 + njobs) (stored-grid fallback of the fused
// loss head); zero n doubles; both only if *cond != 0 (cond may be NULL = always)
int launch_vjp_accumulate(const double* gp, long job0, long njobs, int A, int B, int M, int D, int pairs, const double* gout,
                          double w_diag, double w_off, double* gradX, const unsigned int* cond, cudaStream_t st);
int launch_cond_zero(double* ptr, size_t n, const unsigned int* cond, cudaStream_t st);

// ---- tile forward kernel (skb_tile.cuh): one pair per lane, one strip per warp, W-warp pipelines ---------------
struct TArgs;
// true if skb_sigkernel_fwd should take the tile path: fused kind, scheme S2, GRAM / BATCH pairs, N >= 4, strip shape
// instantiated, and enough (tile, strip) units to fill the GPU
bool tile_applies(int kind, int A, int B, int M, int N, int D, int logd, bool s1, int pairs);
// bytes of the tile path's part of the forward workspace (top-boundary differences, band boundaries, ready counters)
size_t tile_workspace_bytes(int A, int B, int M, int N, int logd, int pairs);
// scale of the prepared rows on the tile path: the exp argument arrives multiplied by 2048 / ln 2 (RBF)
double tile_arg_scale();
// args: Xp, Yp, out, counter (zeroed), A, B, M, N, D, Dp, pairs filled; tile_ws = tile_workspace_bytes() bytes
int launch_tile_forward(int kind, int logd, const KArgs& args, void* tile_ws, cudaStream_t st);
int launch_group_tile_rbf(int rc, int logd, int dp2, const TArgs&, cudaStream_t);
int launch_group_tile_lin(int rc, int logd, int dp2, const TArgs&, cudaStream_t);
// development / tuning knob (process-wide): 0 = never, 1 = whenever the shape is instantiated, -1 = default heuristic
void set_tile_mode(int mode);

#ifdef __CUDACC__
// signal every rank, wait for every rank: lane q of the calling warp takes rank q (called by a full warp once every block
// of the launch has fenced its stores)
__device__ __forceinline__ void rank_barrier(const KArgs& p) {
    __threadfence_system();
    const int q = threadIdx.x & 31;
    if (q < p.n_peer) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.sig_peer[q] + p.sig_rank), "l"(p.sig_epoch) : "memory");
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p.sig_peer[p.sig_rank] + q) : "memory");
        } while (v < p.sig_epoch);
    }
    __syncwarp();
}

#endif

}  // namespace skb
