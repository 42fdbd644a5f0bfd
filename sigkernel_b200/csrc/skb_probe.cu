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

}  // namespace skb

extern "C" int skb_fp64_probe(int op, int blocks, int threads, int iters, double* sink, void* stream) {
    using namespace skb;
    if (blocks <= 0 || threads <= 0 || threads > 256 || iters <= 0) return SKB_ERR_BAD_SHAPE;
    if (!sink) return SKB_ERR_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    if (op == 0) fp64_probe_kernel<0><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 1) fp64_probe_kernel<1><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else if (op == 2) fp64_probe_kernel<2><<<blocks, threads, 0, st>>>(iters, 1.0, sink);
    else return SKB_ERR_BAD_ENUM;
    return check_launch();
}
