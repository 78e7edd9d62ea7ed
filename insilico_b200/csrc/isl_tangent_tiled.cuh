// =============================================================================
// isl_tangent_tiled.cuh -- register-tiled accumulation of the hyperelastic tangent (solid/HyperElastic.hpp:110-179)
//   K[(M,i),(N,k)] = sum_q detJ_q w_q  sum_{J,L} g_M[J] Ceff_q[i,J,k,L] g_N[L]
// for the generic staged kernel.  k_tangent gives one thread one matrix entry: 9 multiply-adds per 15 shared-memory
// loads and quadrature point, i.e. bound by shared-memory bandwidth (Q2 hex: 2.7 M loads per element).  Here a thread
// owns a trial node N and MC test nodes: per point it contracts Ceff with g_N once, T[i,J,k] = sum_L Ceff[i,J,k,L]
// g_N[L] (81 multiply-adds, 84 loads), and then spends 27 multiply-adds per test node on 3 loads; the 9 MC results
// stay in registers over the whole quadrature loop.  Loads per multiply-add drop from 1.7 to 0.4.
// ISL_TANGENT_TILED=1 selects it; NOT YET RUN ON A GPU (written without GPU minutes): the per-thread routine is
// replayed on the host against the defining formula (tests/emu/tangent_tiled_emu.cpp).
// =============================================================================
#pragma once

#if defined(__CUDACC__)
#define ISL_TT_HD __host__ __device__ __forceinline__
#else
#define ISL_TT_HD inline
#endif

// one work item: trial node N, test nodes M0 .. M0+MC-1 (clamped to nt-1; the caller discards the clamped ones).
// Staged arrays of ONE element: Gt [nq][nt][DIM], Gc [nq][nc][DIM], Q [nq][81] with Ceff[i,J,k,L] at ((i*3+J)*3+k)*3+L,
// det [nq]; w [nq] quadrature weights.
template <int DIM, int MC>
ISL_TT_HD void isl_hypel_tile(const double* Gt, const double* Gc, const double* Q, const double* det, const double* w, int nq,
                              int nt, int nc, int N, int M0, double (&acc)[MC][DIM * DIM]) {
    for (int m = 0; m < MC; m++)
        for (int e = 0; e < DIM * DIM; e++) acc[m][e] = 0.;
    for (int q = 0; q < nq; q++) {
        const double* gN = Gc + ((size_t)q * nc + N) * DIM;
        const double* ce = Q + (size_t)q * 81;
        double T[DIM][DIM][DIM];  // [i][J][k]
        for (int i = 0; i < DIM; i++)
            for (int J = 0; J < DIM; J++)
                for (int k = 0; k < DIM; k++) {
                    const double* c4 = ce + ((i * 3 + J) * 3 + k) * 3;
                    double s = c4[0] * gN[0];
                    for (int L = 1; L < DIM; L++) s += c4[L] * gN[L];
                    T[i][J][k] = s;
                }
        const double wd = det[q] * w[q];
        for (int m = 0; m < MC; m++) {
            const int M = (M0 + m < nt) ? M0 + m : nt - 1;
            const double* gM = Gt + ((size_t)q * nt + M) * DIM;
            for (int i = 0; i < DIM; i++)
                for (int k = 0; k < DIM; k++) {
                    double s = gM[0] * T[i][0][k];
                    for (int J = 1; J < DIM; J++) s += gM[J] * T[i][J][k];
                    acc[m][i * DIM + k] += s * wd;
                }
        }
    }
}

#if defined(__CUDACC__)
template <int DIM>
__global__ void __launch_bounds__(256) k_tangent_hypel_tiled(const AsmParams p) {
    constexpr int MC = 6;
    extern __shared__ double smem[];
    const Stage<DIM> s(smem, p);
    const int tid = threadIdx.x, nth = blockDim.x;
    const int nr = p.nt * p.dst, ncl = p.nc * p.dsc;
    const int nchunk = (p.nt + MC - 1) / MC;
    for (int64_t base = (int64_t)blockIdx.x * p.EB; base < p.n_elems; base += (int64_t)gridDim.x * p.EB) {
        const int nb = (int)min((int64_t)p.EB, p.n_elems - base);
        stage_batch<DIM>(p, s, base, nb);
        // effective elasticity per quadrature point (same as k_tangent)
        for (int t = tid; t < nb * p.nq; t += nth) {
            const int eb = t / p.nq;
            double F[3][3], S[3][3], C[6][6];
            deformation_gradient<DIM>(p, s, base + eb, t, F);
            material_eval(p.kernel_id, p.p0, p.p1, F, S, C, true);
            double* ce = s.sQ + (size_t)t * 81;
            for (int i = 0; i < DIM; i++)
                for (int J = 0; J < DIM; J++)
                    for (int k = 0; k < DIM; k++)
                        for (int L = 0; L < DIM; L++) {
                            double r = (i == k ? S[J][L] : 0.);
                            for (int A = 0; A < DIM; A++)
                                for (int B = 0; B < DIM; B++)
                                    r += F[i][A] * C[voigt_idx(A, J)][voigt_idx(B, L)] * F[k][B];
                            ce[((i * 3 + J) * 3 + k) * 3 + L] = r;
                        }
        }
        __syncthreads();
        const int items = nb * p.nc * nchunk;
        for (int t = tid; t < items; t += nth) {
            const int eb = t / (p.nc * nchunk), rem = t % (p.nc * nchunk);
            const int N = rem / nchunk, M0 = (rem % nchunk) * MC;
            double acc[MC][DIM * DIM];
            isl_hypel_tile<DIM, MC>(s.sGt + (size_t)eb * p.nq * p.nt * DIM, s.sGc + (size_t)eb * p.nq * p.nc * DIM,
                                    s.sQ + (size_t)eb * p.nq * 81, s.sDet + (size_t)eb * p.nq, p.w, p.nq, p.nt, p.nc, N, M0, acc);
#pragma unroll
            for (int m = 0; m < MC; m++) {
                const int M = M0 + m;
                if (M < p.nt) {
#pragma unroll
                    for (int i = 0; i < DIM; i++)
#pragma unroll
                        for (int k = 0; k < DIM; k++)
                            scatter_entry(p, base + eb, M * DIM + i, N * DIM + k, nr, ncl, acc[m][i * DIM + k]);
                }
            }
        }
        __syncthreads();
    }
}
#endif
