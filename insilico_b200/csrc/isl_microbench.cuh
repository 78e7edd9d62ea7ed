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

// Measured throughput of FP64 atomic adds without return value (RED.E.ADD.F64) into a large array: the scatter of the
// generic kernels is bounded by it.  pattern 0: consecutive threads add to consecutive entries (fully coalesced sweep);
// 1: every thread adds to three neighbouring entries at a pseudo-random place (what the vector-field kernels do: the
// three components of a node pair, rows of an unstructured numbering); 2: single entries at pseudo-random places.
__global__ void __launch_bounds__(256) k_red_peak(double* a, unsigned long long n, int pattern, int iters) {
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long x = tid * 0x9E3779B97F4A7C15ull + 12345ull;
    for (int it = 0; it < iters; it++) {
        if (pattern == 0) {
            atomicAdd(a + (tid + (unsigned long long)it * nth) % n, 1.0);
        } else {
            x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
            const unsigned long long p = ((x * 0x2545F4914F6CDD1Dull) >> 11) % (n - 3);
            atomicAdd(a + p, 1.0);
            if (pattern == 1) { atomicAdd(a + p + 1, 1.0); atomicAdd(a + p + 2, 1.0); }
        }
    }
}
