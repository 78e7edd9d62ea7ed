// =============================================================================
// isl_microbench.cuh -- measured FP64 peak of the device the engine runs on (included by isl_engine.cu).
// SURVEY 8(d) asks for the FP64-pipe fraction next to the HBM fraction; MEASURED_PEAKS.json has no FP64 figure, so
// the denominator is measured here, on the engine's stream, with CUDA events: independent DFMA chains in registers,
// 64 warps per SM resident, no memory traffic.  2 flops per DFMA.
// =============================================================================
#pragma once

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double b, double c) {
    double a[CHAINS];
#pragma unroll
    for (int j = 0; j < CHAINS; j++) a[j] = (double)(threadIdx.x + j) * 1e-3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int j = 0; j < CHAINS; j++) a[j] = fma(a[j], b, c);
        }
    }
    double s = 0.;
#pragma unroll
    for (int j = 0; j < CHAINS; j++) s += a[j];
    if (s == 123.456) out[0] = s;  // never true for the chosen b, c: keeps the chains alive
}
