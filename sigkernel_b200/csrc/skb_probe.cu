// skb_probe.cu -- register-resident fp64 issue-rate probe: the roofline denominator of the solver.
// MEASURED_PEAKS.json (driver-written) holds HBM and bf16 tensor peaks only; this kernel measures the
// DP-instruction issue rate (DFMA / DADD / DMUL) the stencil is bound by.  Timed by the caller with
// CUDA events on `stream`.  thread-level DP instructions per launch = blocks*threads*iters*16.
#include "skb_host.h"

namespace skb {

template <int OP>
__global__ void __launch_bounds__(256) fp64_probe_kernel(int iters, double seed, double* sink) {
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = seed + (double)(threadIdx.x + i) * 1e-9;
    const double a = 1.0 + seed * 1e-12, b = seed * 1e-13;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) v[i] = fma(v[i], a, b);
            else if (OP == 1) v[i] = __dadd_rn(v[i], b);
            else v[i] = __dmul_rn(v[i], a);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 123.456) sink[0] = s;   // never true; keeps the chains alive
}

// operand-bandwidth variants: OP 3: v = fma(v, w, z) (three distinct register pairs per instruction),
// OP 4: v = fma(v, w, b) (two distinct + one shared), OP 5: v = v + w (two distinct), OP 6: v = fma(a, v, z)
template <int OP>
__global__ void __launch_bounds__(256) fp64_probe3_kernel(int iters, double seed, double* sink) {
    double v[8], w[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[i] = seed + (double)(threadIdx.x + i) * 1e-9;
        w[i] = 1.0 + seed * 1e-12 * (i + 1);
        z[i] = seed * 1e-13 * (i + 1);
    }
    const double a = 1.0 + seed * 1e-12, b = seed * 1e-13;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 3) v[i] = fma(v[i], w[i], z[i]);
                else if (OP == 4) v[i] = fma(v[i], w[i], b);
                else if (OP == 5) v[i] = __dadd_rn(v[i], w[i]);
                else v[i] = fma(a, v[i], z[i]);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i] + w[i] + z[i];
    if (s == 123.456) sink[0] = s;
}

}  // namespace skb

namespace skb {
// mixed-pipe variants: does integer work that fits in the free issue slots slow the fp64 stream down?
// OP 7: 1 DADD (2 distinct) + 1 LOP3 (3 distinct 32-bit register operands) per pair of instructions
// OP 8: 1 DADD + 2 LOP3   OP 9: 1 DADD + 1 IADD with an immediate (1 register operand)
template <int OP>
__global__ void __launch_bounds__(256) fp64_probe_mix_kernel(int iters, double seed, double* sink) {
    double v[8], w[8];
    unsigned a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[i] = seed + (double)(threadIdx.x + i) * 1e-9;
        w[i] = seed * 1e-13 * (i + 1);
        a[i] = threadIdx.x * 7 + i; b[i] = threadIdx.x * 13 + i; c[i] = threadIdx.x * 29 + i;
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[i] = __dadd_rn(v[i], w[i]);
                if (OP == 7 || OP == 8) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));
                if (OP == 8) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(b[i]) : "r"(c[i]), "r"(a[i]));
                if (OP == 9) asm volatile("add.u32 %0, %0, 3;" : "+r"(a[i]));
            }
        }
    }
    double s = 0.0;
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += v[i] + w[i]; t += a[i] + b[i] + c[i]; }
    if (s == 123.456 || t == 0xdeadbeefu) sink[0] = s + t;
}
}  // namespace skb

namespace skb {
// dependent-issue latency: C independent chains of 16 / C DFMAs each per iteration (OP 10: C = 1, 11: C = 2, 12: C = 4,
// 13: C = 8); thread-level DP instructions per launch as for the other variants (16 per iteration).  Run with one
// warp per scheduler: time per instruction = max(issue interval, latency / C).
template <int C>
__global__ void __launch_bounds__(256) fp64_probe_chain_kernel(int iters, double seed, double* sink) {
    double v[C];
#pragma unroll
    for (int i = 0; i < C; ++i) v[i] = seed + (double)(threadIdx.x + i) * 1e-9;
    const double a = 1.0 + seed * 1e-12, b = seed * 1e-13;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 16 / C; ++rep) {
#pragma unroll
            for (int i = 0; i < C; ++i) v[i] = fma(v[i], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += v[i];
    if (s == 123.456) sink[0] = s;
}
}  // namespace skb

namespace skb {
// Does the fp64 tensor-core path (mma.sync m8n8k4 f64) run beside the DFMA pipe?  OP 20: DMMA only (16 per iteration, 4 chains);
// OP 21: 16 DFMA + 4 DMMA per iteration; OP 22: 16 DFMA only (reference for 21).  "Instructions" per launch are counted by the caller.
template <int OP>
__global__ void __launch_bounds__(256) fp64_probe_dmma_kernel(int iters, double seed, double* sink) {
    double c[4][2], v[16];
    const double a = 1.0 + seed * 1e-12, b = seed * 1e-13;
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = seed * i; c[i][1] = seed + i; }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = seed + (double)(threadIdx.x + i) * 1e-9;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (OP == 20) {
#pragma unroll
            for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                v[i] = fma(v[i], a, b);
                if (OP == 21 && (i & 3) == 3)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(c[i >> 2][0]), "+d"(c[i >> 2][1]) : "d"(a), "d"(b));
            }
        }
    }
    double s2 = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s2 += v[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s2 += c[i][0] + c[i][1];
    if (s2 == 123.456) sink[0] = s2;
}
}  // namespace skb

extern "C" int skb_fp64_probe(int op, int blocks, int threads, int iters, double* sink, void* stream) {
    using namespace skb;
    if (blocks <= 0 || threads <= 0 || threads > 256 || iters <= 0) return SKB_ERR_BAD_SHAPE;
    if (!sink) return SKB_ERR_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    if (op == 0) fp64_probe_kernel<0><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 1) fp64_probe_kernel<1><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 2) fp64_probe_kernel<2><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 3) fp64_probe3_kernel<3><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 4) fp64_probe3_kernel<4><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 5) fp64_probe3_kernel<5><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 6) fp64_probe3_kernel<6><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 7) fp64_probe_mix_kernel<7><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 8) fp64_probe_mix_kernel<8><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 9) fp64_probe_mix_kernel<9><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 20) fp64_probe_dmma_kernel<20><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 21) fp64_probe_dmma_kernel<21><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 22) fp64_probe_dmma_kernel<22><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 10) fp64_probe_chain_kernel<1><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 11) fp64_probe_chain_kernel<2><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 12) fp64_probe_chain_kernel<4><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 13) fp64_probe_chain_kernel<8><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else return SKB_ERR_BAD_ENUM;
    return check_launch();
}
