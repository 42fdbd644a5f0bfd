/*
 * oracle/solver.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's CPU Goursat-PDE solvers, used only as
 * the parity checker (tests/, __graft_entry__.smoke(), bench.py cpu_baseline /
 * --impl reference).  Nothing under sigkernel_b200/ may import, link or call it.
 *
 * Restated from (reference = crispitagorico/sigkernel @ 40a5831):
 *   sigkernel/cython_backend.pyx:7-33    sigkernel_cython        -> skb_oracle_solve_batch
 *   sigkernel/cython_backend.pyx:64-119  sigkernel_Gram_cython   -> skb_oracle_solve_gram
 *
 * Arithmetic contract: IEEE double, round-to-nearest, no FMA
 * contraction (build with -ffp-contract=off), inc**2 evaluated as inc*inc,
 * 1./12 folded to a constant:
 *   S2 (default):      u11 = (u10 + u01) * ((1 + 0.5 g) + (1/12)(g g)) - u00 * (1 - (1/12)(g g))
 *   S1 (_naive_solver): u11 = (u10 + u01) * (1 + 0.5 g) - u00
 * Pinned bit-for-bit against the compiled reference (oracle/_ref) by
 * tests/test_oracle.py::test_c_restatement_bitwise_vs_ref and against the
 * committed golden fixtures (tests/golden/).
 *
 * Layout: inc is row-major (pairs, MM, NN); the solution grid is row-major
 * (pairs, MM+1, NN+1) with u[.,0,:] = u[.,:,0] = 1.
 */
#include <stddef.h>

static void solve_one(const double *g, int MM, int NN, int naive, double *u)
{
    const size_t ld = (size_t)NN + 1;
    const double twelfth = 1. / 12;
    for (int j = 0; j <= NN; ++j) u[j] = 1.;
    for (int i = 0; i <= MM; ++i) u[(size_t)i * ld] = 1.;
    for (int i = 0; i < MM; ++i) {
        const double *grow = g + (size_t)i * NN;
        const double *up = u + (size_t)i * ld;       /* node row i   */
        double *cur = u + (size_t)(i + 1) * ld;      /* node row i+1 */
        for (int j = 0; j < NN; ++j) {
            const double gij = grow[j];
            if (naive) {
                cur[j + 1] = (cur[j] + up[j + 1]) * (1. + 0.5 * gij) - up[j];
            } else {
                cur[j + 1] = (cur[j] + up[j + 1]) * (1. + 0.5 * gij + twelfth * (gij * gij))
                             - up[j] * (1. - twelfth * (gij * gij));
            }
        }
    }
}

/* cython_backend.pyx:7-33 */
void skb_oracle_solve_batch(const double *inc, int A, int MM, int NN, int naive, double *K)
{
    const size_t cells = (size_t)MM * NN, nodes = (size_t)(MM + 1) * (NN + 1);
    for (int a = 0; a < A; ++a)
        solve_one(inc + a * cells, MM, NN, naive, K + a * nodes);
}

/* Diagonal pair (l,l) of the symmetric Gram: the reference writes the mirrored node
 * K[l,l,j+1,i+1] = K[l,l,i+1,j+1] WHILE it sweeps (cython_backend.pyx:97 with m == l), so
 * node (i,i+1) is replaced by the freshly computed (i+1,i) before cell (i,i) reads it.
 * With increments that are symmetric only up to rounding this changes the last bits
 * (relative ~1e-14 on entries of size 1e6); restated here so the oracle stays bit-faithful. */
static void solve_one_selfmirror(const double *g, int MM, int naive, double *u)
{
    const size_t ld = (size_t)MM + 1;
    const double twelfth = 1. / 12;
    for (int j = 0; j <= MM; ++j) u[j] = 1.;
    for (int i = 0; i <= MM; ++i) u[(size_t)i * ld] = 1.;
    for (int i = 0; i < MM; ++i) {
        for (int j = 0; j < MM; ++j) {
            const double gij = g[(size_t)i * MM + j];
            const double u10 = u[(size_t)(i + 1) * ld + j], u01 = u[(size_t)i * ld + j + 1];
            const double u00 = u[(size_t)i * ld + j];
            double v;
            if (naive)
                v = (u10 + u01) * (1. + 0.5 * gij) - u00;
            else
                v = (u10 + u01) * (1. + 0.5 * gij + twelfth * (gij * gij))
                    - u00 * (1. - twelfth * (gij * gij));
            u[(size_t)(i + 1) * ld + j + 1] = v;
            u[(size_t)(j + 1) * ld + i + 1] = v;
        }
    }
}

/* cython_backend.pyx:64-119.  sym!=0 requires A==B, MM==NN (the reference
 * mirrors the transposed grid of pair (l,m) into pair (m,l), :97). */
void skb_oracle_solve_gram(const double *inc, int A, int B, int MM, int NN, int sym, int naive,
                           double *K)
{
    const size_t cells = (size_t)MM * NN, nodes = (size_t)(MM + 1) * (NN + 1);
    if (!sym) {
        for (size_t p = 0; p < (size_t)A * B; ++p)
            solve_one(inc + p * cells, MM, NN, naive, K + p * nodes);
        return;
    }
    for (int l = 0; l < A; ++l) {
        for (int m = l; m < A; ++m) {
            double *ulm = K + ((size_t)l * B + m) * nodes;
            double *uml = K + ((size_t)m * B + l) * nodes;
            if (m == l) {
                solve_one_selfmirror(inc + ((size_t)l * B + m) * cells, MM, naive, ulm);
                continue;
            }
            solve_one(inc + ((size_t)l * B + m) * cells, MM, NN, naive, ulm);
            for (int i = 0; i <= MM; ++i)
                for (int j = 0; j <= NN; ++j)
                    uml[(size_t)j * (MM + 1) + i] = ulm[(size_t)i * (NN + 1) + j];
        }
    }
}

/* Corner-only variant for the CPU baseline at large shapes: same arithmetic,
 * two live rows instead of the full grid (so cfg3/cfg5 samples fit in RAM).
 * out[p] = u[MM,NN] of pair p.  Optional OpenMP over pairs (bench only). */
void skb_oracle_solve_gram_corner(const double *inc, long pairs, int MM, int NN, int naive,
                                  double *out, double *scratch /* 2*(NN+1) per thread */)
{
    const size_t cells = (size_t)MM * NN;
    const double twelfth = 1. / 12;
    for (long p = 0; p < pairs; ++p) {
        const double *g = inc + (size_t)p * cells;
        double *up = scratch, *cur = scratch + (NN + 1);
        for (int j = 0; j <= NN; ++j) up[j] = 1.;
        for (int i = 0; i < MM; ++i) {
            cur[0] = 1.;
            for (int j = 0; j < NN; ++j) {
                const double gij = g[(size_t)i * NN + j];
                if (naive)
                    cur[j + 1] = (cur[j] + up[j + 1]) * (1. + 0.5 * gij) - up[j];
                else
                    cur[j + 1] = (cur[j] + up[j + 1]) * (1. + 0.5 * gij + twelfth * (gij * gij))
                                 - up[j] * (1. - twelfth * (gij * gij));
            }
            double *t = up; up = cur; cur = t;
        }
        out[p] = up[NN];
    }
}

/* sigkernel/cuda_backend.py:165-223  sigkernel_derivatives_Gram_cuda -> skb_oracle_solve_derivatives
 * (the reference's CPU branch of k_kgrad is broken, sigkernel.py:588 vs cython_backend.pyx:176, so the
 * restatement follows the Numba kernel; pinned against that kernel run under Numba's CUDA simulator,
 * tests/golden/make_golden.py).  Row-major sweep instead of the kernel's anti-diagonal order: every cell
 * reads only finished cells, the arithmetic per cell is the kernel's, statement by statement.
 * inc, incd, incdd: (pairs, MM, NN); out: (pairs, 3) = K, K_diff, K_diffdiff at node (MM, NN);
 * work: 3 * (MM+1) * (NN+1) doubles. */
void skb_oracle_solve_derivatives(const double *inc, const double *incd, const double *incdd, long pairs,
                                  int MM, int NN, double *out, double *work)
{
    const size_t ld = (size_t)NN + 1, nodes = (size_t)(MM + 1) * ld, cells = (size_t)MM * NN;
    double *K = work, *Kd = work + nodes, *Kdd = work + 2 * nodes;
    for (long p = 0; p < pairs; ++p) {
        const double *g = inc + p * cells, *gd = incd + p * cells, *gdd = incdd + p * cells;
        for (size_t k = 0; k < nodes; ++k) { K[k] = 0.; Kd[k] = 0.; Kdd[k] = 0.; }
        for (int j = 0; j <= NN; ++j) K[j] = 1.;
        for (int i = 0; i <= MM; ++i) K[(size_t)i * ld] = 1.;
        for (int i = 1; i <= MM; ++i) {
            for (int j = 1; j <= NN; ++j) {
                const double in = g[(size_t)(i - 1) * NN + j - 1];
                const double ind = gd[(size_t)(i - 1) * NN + j - 1];
                const double indd = gdd[(size_t)(i - 1) * NN + j - 1];
                const size_t c11 = (size_t)i * ld + j, c01 = c11 - ld, c10 = c11 - 1, c00 = c01 - 1;
                const double k01 = K[c01], k10 = K[c10], k00 = K[c00];
                const double k01d = Kd[c01], k10d = Kd[c10], k00d = Kd[c00];
                const double k01dd = Kdd[c01], k10dd = Kdd[c10], k00dd = Kdd[c00];
                const double k11 = (k01 + k10) * (1. + 0.5 * in + (1. / 12) * (in * in)) - k00 * (1. - (1. / 12) * (in * in));
                K[c11] = k11;
                const double f1 = k00 * ind + k00d * in;
                const double f2 = k01 * ind + k01d * in;
                const double f3 = k10 * ind + k10d * in;
                const double f4 = k11 * ind + (k01d + k10d - k00d + f1) * in;
                const double k11d = k01d + k10d - k00d + 0.25 * (f1 + f2 + f3 + f4);
                Kd[c11] = k11d;
                const double g1 = k00 * indd + 2. * k00d * ind + k00dd * in;
                const double g2 = k01 * indd + 2. * k01d * ind + k01dd * in;
                const double g3 = k10 * indd + 2. * k10d * ind + k10dd * in;
                const double g4 = k11 * indd + 2. * k11d * ind + (k01dd + k10dd - k00dd + g1) * in;
                Kdd[c11] = k01dd + k10dd - k00dd + 0.25 * (g1 + g2 + g3 + g4);
            }
        }
        out[3 * p + 0] = K[nodes - 1];
        out[3 * p + 1] = Kd[nodes - 1];
        out[3 * p + 2] = Kdd[nodes - 1];
    }
}
