/*
 * sigkernel_b200.h -- C ABI of the B200-native signature-kernel PDE solver.
 *
 * This is the drop-in boundary for the hot path of crispitagorico/sigkernel
 * (reference @ 40a5831; all file:line citations are relative to the reference tree):
 * plain pointers and sizes, no torch types.  Every pointer named "device" is a CUDA
 * device pointer to contiguous row-major memory owned by the CALLER; nothing is
 * allocated inside the library; every launch is asynchronous on the `stream` argument
 * (a cudaStream_t passed as void*; NULL = legacy default stream).  The compute entry points keep
 * no state between calls and may be called from several threads at once; the only process-wide
 * state is the three tuning / measurement knobs below (skb_set_warps_per_sm, skb_set_tile_mode:
 * plain ints read at launch time; skb_set_profile_events: per calling thread).
 *
 * Notation: A, B batch sizes; M, N path lengths (points); D path dimension; d dyadic
 * order; MM = (M-1) << d, NN = (N-1) << d fine cells per axis.
 *
 * Return value of every int function: SKB_OK (0) or a negative SKB_ERR_* code
 * (skb_error_string() names it).  The reference has no error reporting on its GPU path
 * beyond Python asserts (sigkernel/sigkernel.py:222,368); the thread-per-row limit
 * max(MM,NN) < 1024 of those asserts does not exist here.
 */
#ifndef SIGKERNEL_B200_H
#define SIGKERNEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums (plain ints in the ABI) ------------------------------------------------ */

/* static kernel evaluated on the fly (sigkernel/static_kernels.py:11-73) */
#define SKB_STATIC_LINEAR 0 /* k(x,y) = param * <x,y>; param = scale^2 for batch_kernel (:24), 1 for Gram_matrix (:33) */
#define SKB_STATIC_RBF    1 /* k(x,y) = exp(-|x-y|^2 / param); param = sigma (:42-73) */

/* finite-difference scheme of the Goursat PDE (sigkernel/cython_backend.pyx:27,30,91,94,114,116) */
#define SKB_SCHEME_S2 0 /* default: (u10+u01)(1 + g/2 + g^2/12) - u00 (1 - g^2/12) */
#define SKB_SCHEME_S1 1 /* _naive_solver=True: (u10+u01)(1 + g/2) - u00 */

/* which (a,b) pairs are solved */
#define SKB_PAIRS_GRAM  0 /* all A*B pairs, out[a*B+b]           (sigkernel_Gram_cuda, cuda_backend.py:121-160) */
#define SKB_PAIRS_BATCH 1 /* pairs (a,a), requires A == B, out[a] (sigkernel_cuda, cuda_backend.py:6-49) */
#define SKB_PAIRS_SYM   2 /* X is Y: solve a<=b, mirror into out[b*A+a] (cython_backend.pyx:76-97) */

/* arithmetic mode of the stencil */
#define SKB_ARITH_FMA   0 /* fused multiply-add form, 3 fp64 instructions per cell (default, fastest) */
#define SKB_ARITH_EXACT 1 /* the reference's operation order with no FMA contraction: bit-identical
                             to cython_backend.pyx given identical increments (4 instructions per cell) */

/* element type of host-facing path / output buffers */
#define SKB_F64 0
#define SKB_F32 1

#define SKB_OK               0
#define SKB_ERR_BAD_SHAPE   -1 /* non-positive size, M or N < 2, BATCH/SYM with A != B ... */
#define SKB_ERR_BAD_ENUM    -2 /* unknown static kind / scheme / pairs / arith / dtype */
#define SKB_ERR_WORKSPACE   -3 /* workspace pointer NULL or smaller than skb_*_workspace_bytes() */
#define SKB_ERR_UNSUPPORTED -4 /* shape outside what this build instantiates (see skb_error_string) */
#define SKB_ERR_CUDA        -5 /* a CUDA call failed; skb_last_cuda_error() has the cudaError_t */
#define SKB_ERR_NULL        -6 /* a required pointer is NULL */

const char* skb_error_string(int code);
int         skb_last_cuda_error(void);     /* cudaError_t of the last SKB_ERR_CUDA on this thread */
int         skb_version(void);             /* ABI version, bumped on any signature change */

/* Tuning knob (process-wide, default 0 = automatic): resident solver warps per SM. */
void skb_set_warps_per_sm(int warps);

/* Tuning knob (process-wide, default -1): which forward kernel serves skb_sigkernel_fwd for the fused static
 * kinds.  1 = the experimental tile kernel (one pair per lane, one strip per warp, skb_tile.cuh) whenever the shape
 * is instantiated (dyadic order 1..3, len_y >= 16, GRAM / BATCH pairs); 0 or -1 = fwd5_kernel / solver_kernel
 * (the tile kernel is slower than fwd5_kernel at every BASELINE config so far). */
void skb_set_tile_mode(int mode);

/* Tuning / test knob (process-wide, default -1): which kernels serve the backward entry points.  -1 (or 1) = adjoint by
 * reconstruction (32 lanes per pair) with the stored-grid kernels queued as a device-side fallback, 0 = stored-grid
 * kernels only (round-1 behaviour), 2 = reconstruction with 16 lanes per pair where instantiated, 3 = as -1 but without
 * the unordered-pair sweep of Gram(X, X) (skb_adjoint_sym_supported returns 0). */
void skb_set_adjoint_mode(int mode);

/* Tuning / test knob (process-wide, default -1): skb_sigkernel_derivatives_from_static.  -1 (or 1) = the streaming kernel
 * (one warp per pair, FMA-contracted: agrees with the reference's arithmetic to rounding) wherever (M - 1) 2^d <= 256 and
 * d <= 3, the diagonal kernel elsewhere; 0 = the diagonal kernel always (the reference's operation order, bit for bit). */
void skb_set_deriv_mode(int mode);

/* Measurement hook (per calling thread): when both are non-NULL, every solver launch made by this thread records
 * `start` immediately before and `stop` immediately after the solver kernel on the launch stream
 * (cudaEvent_t passed as void*), so a caller can time the dominant kernel alone, without the
 * path-preparation kernels around it.  Pass NULLs to disable. */
void skb_set_profile_events(void* start_event, void* stop_event);

/* Diagnostic used by bench.py for the roofline denominator: launches a register-resident chain of
 * fp64 instructions (op 0 = DFMA, 1 = DADD, 2 = DMUL); blocks*threads*iters*16 thread-level DP
 * instructions per launch, timed by the caller with CUDA events on `stream`.  threads <= 256. */
int skb_fp64_probe(int op, int blocks, int threads, int iters, double* sink, void* stream);

/* Which kernel family a call would use (host-side dispatch only, no GPU work; for tests and diagnostics):
 *   skb_forward_plan  (skb_sigkernel_fwd):      0 = generic row-band sweep of the fine grid (any shape),
 *                                               1 = register-resident solver_kernel (v4),
 *                                               4 = fwd5_kernel with 16 lanes per pair (two pairs per warp),
 *                                               5 / 6 / 7 = fwd5_kernel with 1 / 2 / 4 warps per pair;
 *   skb_adjoint_plan  (skb_sigkernel_fwd_bwd): 1 = solver_kernel store / reversed modes (v4),
 *                                               5 = fwd5_kernel store / reversed modes,
 *                                               6 = adjoint by reconstruction (fwd5_kernel MODE_FWD_EMIT + MODE_REV_RECON;
 *                                                   no stored grid, (len_x - 1) 2^d <= 1024 at dyadic order <= 2),
 *                                               7 = every other length: the reference's algebra on materialised grids
 *                                                   (forward grid, reversed grid, pooled product; skb_generic_adj.cu) --
 *                                                   no shape is refused, skb_sigkernel_fwd_bwd only.
 * Negative values are SKB_ERR_* codes for bad arguments. */
int skb_forward_plan(int M, int N, int D, int dyadic_order, int static_kind, int scheme);
int skb_adjoint_plan(int M, int N, int D, int dyadic_order, int static_kind, int scheme);
/* 1 if skb_sigkernel_bwd_vjp takes pairs = SKB_PAIRS_SYM for paths of this length (Gram(X, X): one reversed sweep per
 * unordered pair yields the gradient w.r.t. both paths), else 0. */
int skb_adjoint_sym_supported(int M, int D, int dyadic_order, int static_kind, int scheme);

/* ---- workspaces -------------------------------------------------------------------
 * Every compute entry point takes a caller-owned device scratch buffer (256-byte aligned) that
 * holds the job-queue counter, the prepared paths and (backward) the forward solution grids.
 * The *_workspace_bytes functions return the size to allocate; 0 means bad arguments.
 * Shapes the register-resident kernels do not cover (ceil(M/32) rounded up to a power of two, times
 * 2^dyadic_order, > 32) are served by a generic row-band sweep of the fine grid that needs a larger
 * workspace (<= ~1 GiB; the pairs are processed in chunks that fit). */
size_t skb_fwd_workspace_bytes(int A, int B, int M, int N, int D, int dyadic_order, int pairs);
/* from_static (A,B,M,N as passed there) / solve_increments (A = B = P, M = MM+1, N = NN+1, dyadic_order 0,
 * pairs = SKB_PAIRS_BATCH) */
size_t skb_aux_workspace_bytes(int A, int B, int M, int N, int dyadic_order, int pairs);
/* skb_sigkernel_sensitivity_from_static (any length: materialised grids beyond the register-resident kernels) */
size_t skb_sensitivity_workspace_bytes(int A, int B, int M, int N, int dyadic_order, int pairs);
/* Recommended size for skb_sigkernel_fwd_bwd / skb_sigkernel_sensitivity_from_static: prepared paths, the boundary
 * context of the adjoint by reconstruction, and room for the forward grids of the stored-grid kernels (all pairs, capped
 * at 8 GiB; 1 GiB when they are only the fallback of the reconstruction).  Any size >= the fixed part + ONE pair's grid
 * is accepted: the pairs are then processed in chunks that fit. */
size_t skb_bwd_workspace_bytes(int A, int B, int M, int N, int D, int dyadic_order, int pairs);
/* Boundary context written by skb_sigkernel_fwd_ctx and read by skb_sigkernel_bwd_vjp: the last row and the last column of
 * every pair's solution grid ((NN + 1) + (MM + 1) doubles per pair, rounded up to multiples of 4). */
size_t skb_ctx_bytes(int A, int B, int M, int N, int dyadic_order, int pairs);
size_t skb_bwd_vjp_workspace_bytes(int A, int B, int M, int N, int D, int dyadic_order, int pairs);

/* ---- forward: fused static kernel + increments + dyadic refinement + PDE solve ----- */

/*
 * out[pairs] = signature kernel k(X_a, Y_b) for the pair set `pairs`.
 * Replaces, in one launch sequence and without materialising any (A,B,MM,NN) tensor:
 *   static_kernel.Gram_matrix / batch_kernel      static_kernels.py:17-33, 42-73
 *   second difference + tile()                     sigkernel.py:217-218, 362-364, 607-613
 *   sigkernel_Gram_cuda / sigkernel_cuda launch    sigkernel.py:224-234, 370-382; cuda_backend.py:6-49, 121-160
 *   (= sigkernel_Gram_cython / sigkernel_cython    cython_backend.pyx:7-33, 64-119 on the CPU branch)
 * and the final slice K[..., -1, -1] (sigkernel.py:253, 401).
 *
 * X (A,M,D), Y (B,N,D): device, element type `io_dtype`.  out: device, fp64, A*B entries
 * (GRAM, SYM) or A entries (BATCH).  Arithmetic is fp64 regardless of io_dtype.
 * `arith` must be SKB_ARITH_FMA here (the fused static kernel is not bit-comparable anyway).
 */
int skb_sigkernel_fwd(const void* X, const void* Y, int io_dtype,
                      int A, int B, int M, int N, int D, int dyadic_order,
                      int static_kind, double static_param, int scheme, int pairs, int arith,
                      double* out, void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same forward with the result written to SEVERAL destinations: out_peers is a HOST array of n_peers (<= 8) device
 * pointers, each laid out like `out`; every k(X_a, Y_b) is stored to all of them.  This is the multi-GPU gather of
 * sigkernel_b200.distributed without a collective: each rank passes, for every rank q, the address of ITS OWN row block
 * inside q's copy of G (peer memory mapped through NVLink, e.g. torch symmetric memory), so the Gram matrix assembles
 * itself on every rank while the solver runs; a barrier across the ranks afterwards is all that remains.
 * GRAM / BATCH pairs, shapes served by fwd5_kernel (skb_forward_plan >= 4); SKB_ERR_UNSUPPORTED otherwise.
 */
int skb_sigkernel_fwd_peers(const void* X, const void* Y, int io_dtype,
                            int A, int B, int M, int N, int D, int dyadic_order,
                            int static_kind, double static_param, int scheme, int pairs,
                            double* const* out_peers, int n_peers,
                            void* workspace, size_t workspace_bytes, void* stream);

/*
 * The static kernel matrix itself: Ks[pair][i][j] = kappa(X_a[i], Y_b[j]) (fp64; (A,B,M,N) for GRAM, (A,M,N) for BATCH) for
 * the two built-in kernels (static_kernels.py:17-33 Linear, :42-73 RBF) -- what the reference's plugin interface
 * (`Gram_matrix` / `batch_kernel`) returns, in one pass over the output.  Used by the derivative path, which needs the
 * matrices of three perturbed copies of X (sigkernel.py:524-539).  Workspace: skb_fwd_workspace_bytes (prepared paths).
 */
int skb_static_gram(const void* X, const void* Y, int io_dtype, int A, int B, int M, int N, int D,
                    int static_kind, double static_param, int pairs, double* Ks,
                    void* workspace, size_t workspace_bytes, void* stream);

/*
 * One slice [job_lo, job_hi) of the pair enumeration of `pairs` -- a rank's share of a sharded Gram matrix.  GRAM: job =
 * a * B + b; BATCH: job = a; SYM: the pairs a <= b, row by row (job = a * A - a (a - 1) / 2 + (b - a)) -- equal slices of
 * the SYM enumeration are equal amounts of work, which contiguous row blocks of a symmetric matrix are not.
 * Results go to `out` (n_peers == 0; laid out like the full matrix, entries of other jobs are left alone) or to every
 * pointer of out_peers (n_peers > 0, as skb_sigkernel_fwd_peers; each points to the START of a rank's full copy); SYM
 * writes both mirror entries.  Shapes served by fwd5_kernel (skb_forward_plan >= 4); SKB_ERR_UNSUPPORTED otherwise.
 * sig_peers (may be NULL; needs n_peers > 0): HOST array of n_peers device pointers, rank q's array of n_peers 64-bit
 * signal slots (peer memory, zero before the first call).  The solver kernel then ENDS with the barrier across the ranks:
 * its last block stores sig_epoch (> 0, increasing from call to call) to slot my_rank of every rank's array and waits until
 * all slots of its own array have reached sig_epoch -- when the call's stream work is done, every rank's results are in
 * this rank's copy, with no barrier kernel or collective behind the solve.  Every rank of the group must make the call.
 */
int skb_sigkernel_fwd_range(const void* X, const void* Y, int io_dtype,
                            int A, int B, int M, int N, int D, int dyadic_order,
                            int static_kind, double static_param, int scheme, int pairs,
                            long job_lo, long job_hi, double* out, double* const* out_peers, int n_peers,
                            unsigned long long* const* sig_peers, int my_rank, unsigned long long sig_epoch,
                            void* workspace, size_t workspace_bytes, void* stream);

/*
 * Plugin path: the caller evaluated an arbitrary static kernel itself
 * (any object with Gram_matrix / batch_kernel, static_kernels.py:75-206) and passes the
 * COARSE matrix Ks: (A,B,M,N) for GRAM/SYM, (A,M,N) for BATCH, device fp64.  Second
 * difference and dyadic refinement happen on the fly.  With SKB_ARITH_EXACT the result is
 * bit-identical to the reference CPU branch fed the same Ks.
 */
int skb_sigkernel_fwd_from_static(const double* Ks, int A, int B, int M, int N, int dyadic_order,
                                  int scheme, int pairs, int arith,
                                  double* out, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Operator-level mirror of the reference's L3->L1 call (sigkernel.py:378-380 /
 * cython_backend.pyx:64): the caller built the fine increment tensor inc (P,MM,NN) itself;
 * out[p] = u[MM,NN].  With SKB_ARITH_EXACT bit-identical to sigkernel_Gram_cython(...)[..,-1,-1].
 */
int skb_sigkernel_solve_increments(const double* inc, long P, int MM, int NN, int scheme, int arith,
                                   double* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- backward: adjoint (reversed) PDE -> per-point gradients ---------------------- */

/*
 * Forward value AND the reference's `grad_points` in one call (the reference computes the
 * backward eagerly inside forward when X.requires_grad, sigkernel.py:397-399):
 *   out[pairs]                 as skb_sigkernel_fwd
 *   grad_points (pairs, M, D)  d k(X_a,Y_b) / d X_a[p,:]  as defined by prep_backward /
 *                              _SigKernel.backward (sigkernel.py:419-502, 256-343): reversed PDE,
 *                              GG = u * u_rev, sum of GG * d inc / d x over the cells touching point p,
 *                              with the ANALYTIC static-kernel derivative in place of the reference's
 *                              h = 1e-9 one-sided finite difference (agrees with it to its own noise floor,
 *                              SURVEY.md 8(a)).
 * `pairs` is SKB_PAIRS_GRAM, SKB_PAIRS_BATCH or SKB_PAIRS_SYM (Y = X: out (A, A) and grad_points (A, A, M, D) are the FULL
 * tensors -- a symmetric Gram needs every (a, b) gradient -- but where the unordered-pair sweep covers the shape
 * (skb_adjoint_sym_supported) they come from one forward solve and one reversed sweep per pair a <= b: the sweep of (a, b)
 * yields d k(X_a,X_b) / d X_a and, from the same sensitivities, d k(X_b,X_a) / d X_b; cython_backend.pyx:76-97 exploits the
 * symmetry in the forward only).
 */
int skb_sigkernel_fwd_bwd(const void* X, const void* Y, int io_dtype,
                          int A, int B, int M, int N, int D, int dyadic_order,
                          int static_kind, double static_param, int scheme, int pairs,
                          double* out, double* grad_points,
                          void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same backward split in two, for callers that learn d loss / d K only later (torch.autograd's backward) and
 * never want the (pairs, M, D) tensor:
 *
 * skb_sigkernel_fwd_ctx   out[pairs] as skb_sigkernel_fwd (GRAM, BATCH or SYM) plus the boundary context `ctx`
 *                         (skb_ctx_bytes): all the reversed sweep needs from the forward solve.  Workspace:
 *                         skb_fwd_workspace_bytes.  SKB_ERR_UNSUPPORTED if the shape is outside the reconstruction
 *                         kernels (see skb_adjoint_plan == 6): use skb_sigkernel_fwd_bwd then.
 * skb_sigkernel_bwd_vjp   reversed sweep over every ordered pair of `pairs` (GRAM or BATCH; `ctx_pairs` names the pair
 *                         set of the forward call: SYM is allowed with GRAM here -- the pair (a, b), a > b, reads the
 *                         transposed grid of (b, a)), contracted on the fly with d loss / d K:
 *                             g[a, m, :] = sum_b coef(a, b) * d k(X_a, Y_b) / d X_a[m, :]
 *                             gradX = (accumulate ? gradX : 0) + out_scale * (out_scale_dev ? *out_scale_dev : 1) * g
 *                         coef = grad_out[pair] if grad_out != NULL, else (a == b ? w_diag : w_off) -- the closed forms of
 *                         the MMD and the scoring rules (sigkernel.py:146-197), replacing sigkernel.py:405-416; out_scale
 *                         carries the reference's factor 2 for Gram(X, X) (sigkernel.py:410-412), out_scale_dev (device,
 *                         may be NULL) the upstream gradient of a scalar loss.  g is summed with fp64 atomics (the order
 *                         over b varies from run to run).  grad_points (pairs, M, D) is also written if not NULL; one of
 *                         gradX (A, M, D) and grad_points must be given.  pairs = SKB_PAIRS_SYM (Y = X, ctx_pairs = SYM, gradX
 *                         only; skb_adjoint_sym_supported): the SAME sum g[a] = sum_b coef(a, b) d k(X_a, X_b) / d X_a from ONE
 *                         sweep per unordered pair a <= b -- the sweep of (a, b) also contracts its sensitivities with the rows
 *                         of X_a per node column, which is d k(X_b, X_a) / d X_b, the term of the ordered pair (b, a)
 *                         (cython_backend.pyx:76-97 solves the triangle only in the forward; here the backward does too):
 *                         half the sweeps of the GRAM call it replaces, same weights, same out_scale.  Workspace: skb_bwd_vjp_workspace_bytes (smaller is accepted: without room for
 *                         one pair's grid + gradients the stored-grid fallback is not queued; the flag word at byte 64 of
 *                         the workspace is then the caller's to check -- non-zero = a rebuilt grid missed its boundary
 *                         by more than 1e-10 and the result should be recomputed with skb_sigkernel_fwd_bwd).
 */
int skb_sigkernel_fwd_ctx(const void* X, const void* Y, int io_dtype,
                          int A, int B, int M, int N, int D, int dyadic_order,
                          int static_kind, double static_param, int scheme, int pairs,
                          double* out, void* ctx, size_t ctx_bytes,
                          void* workspace, size_t workspace_bytes, void* stream);
int skb_sigkernel_bwd_vjp(const void* X, const void* Y, int io_dtype,
                          int A, int B, int M, int N, int D, int dyadic_order,
                          int static_kind, double static_param, int scheme, int pairs,
                          const void* ctx, int ctx_pairs, const double* grad_out, double w_diag, double w_off,
                          double out_scale, const double* out_scale_dev, int accumulate,
                          double* gradX, double* grad_points,
                          void* workspace, size_t workspace_bytes, void* stream);

/* acc[0] = (accumulate ? acc[0] : 0) + sum_{a,b} w(a,b) G[a,b], w = w_diag on the diagonal and w_off elsewhere; G is the
 * row-major (A, B) output of a GRAM / SYM call, or the A-vector of a BATCH call (every entry weighted w_diag): the
 * reductions of compute_mmd / compute_distance / compute_scoring_rule (sigkernel.py:130-197), one launch per Gram. */
int skb_gram_weighted_sum(const double* G, int A, int B, int pairs, double w_diag, double w_off, double* acc, int accumulate,
                          void* stream);

/*
 * Plugin path of the backward: coarse sensitivities
 *   S[pair, i, j] = 4^-d * sum over the fine cells (p,q) of coarse cell (i,j) of u[p,q] * u_rev[p+1,q+1]
 * (pairs, M-1, N-1), from the coarse static matrix Ks.  The caller contracts S with its own
 * d inc / d x (finite-difference Gram_matrix(X+h e_d, Y) exactly as sigkernel.py:473-487 does).
 * `pairs` is SKB_PAIRS_GRAM or SKB_PAIRS_BATCH.
 */
int skb_sigkernel_sensitivity_from_static(const double* Ks, int A, int B, int M, int N, int dyadic_order,
                                          int scheme, int pairs, double* out, double* S,
                                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- kernel and its first / second directional derivative along gamma ----------------------
 *
 * Replaces k_kgrad's solve (sigkernel.py:504-593) and sigkernel_derivatives_Gram_cuda
 * (cuda_backend.py:165-223) behind SigKernel.compute_kernel_and_derivatives_Gram (sigkernel.py:43-89).
 * The caller evaluates the static kernel three times, exactly as the reference does (sigkernel.py:524-539):
 *   K0 = Gram_matrix(X, Y), K1 = Gram_matrix(X + eps*gamma, Y), K2 = Gram_matrix(X + 2*eps*gamma, Y),
 * each (A,B,M,N) device fp64, and passes the COARSE matrices; the finite differences in eps, the second
 * differences, the dyadic refinement and the three coupled stencils run on the device.
 *   out3 (A*B, 3): k, k_gamma, k_gamma_gamma at the end point, interleaved per pair.
 * workspace: skb_deriv_workspace_bytes(A, B, M, N).  Limit: 9 * (MM + 1) doubles of shared memory per pair
 * (MM = (M-1) << dyadic_order <= 2843).
 */
size_t skb_deriv_workspace_bytes(int A, int B, int M, int N);
int skb_sigkernel_derivatives_from_static(const double* K0, const double* K1, const double* K2,
                                          int A, int B, int M, int N, int dyadic_order, double eps,
                                          double* out3, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIGKERNEL_B200_H */
