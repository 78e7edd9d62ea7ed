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
            deformation_gradient<DIM>(p, s, p.eid(base + eb), t, F);
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
                            scatter_entry(p, p.eid(base + eb), M * DIM + i, N * DIM + k, nr, ncl, acc[m][i * DIM + k]);
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// k_tangent_hypel_sym<MC>: the hyperelastic tangent with
//   * the effective elasticity per quadrature point formed by ALL threads in two contraction steps,
//       D[A,J,k,L] = sum_B C[V(A,J),V(B,L)] F[k,B],   Ceff[i,J,k,L] = delta_ik S[J,L] + sum_A F[i,A] D[A,J,k,L]
//     (2 x 243 multiply-adds per point, one thread per (point, J, k, L); k_tangent / the tiled kernel above give a
//     whole point to one thread, 2187 x 2 multiplies, while the other threads of the CTA wait),
//   * symmetry: Ceff[i,J,k,L] = Ceff[k,L,i,J] (S symmetric, C with major symmetry), hence K[(M,i),(N,k)] =
//     K[(N,k),(M,i)]: only node pairs M <= N are integrated and every block is scattered together with its transpose
//     (Q2 hex: 378 instead of 729 node pairs),
//   * register tiles: a thread owns trial node N and up to MC test nodes M <= N; per point it contracts Ceff with
//     det J w g_N once (81 multiply-adds, the 81 numbers are the same for all lanes of a warp working on one element:
//     broadcast 16-byte loads) and spends 27 multiply-adds per test node on 3 loads; 9 MC results stay in registers.
// Multiply-adds per Q2-hex element and point: 75 x 81 + 378 x 27 = 16.3 k (k_tangent: 59 k, the tiled kernel: 32.8 k;
// the algorithmic count of SURVEY 8(d) without symmetry: 21.9 k).
// Shared memory per element (doubles): X npe*3 | U nt*3 | contra nq*9 | det nq | G nq*nt*3 | Ceff nq*82 | F,S,C nq*54.
constexpr int HS_QSTRIDE = 82;   // 81 numbers of Ceff per point, padded to a 16-byte multiple

struct HypelSymLayout {
    int per_elem;   // doubles
    int oX, oU, oCon, oDet, oG, oQ, oM;
    __host__ __device__ HypelSymLayout(int npe, int nq, int nt) {
        oX = 0; oU = oX + npe * 3; oCon = oU + nt * 3; oDet = oCon + nq * 9; oG = oDet + nq; oG += (oG & 1);
        oQ = oG + nq * nt * 3; oQ += (oQ & 1); oM = oQ + nq * HS_QSTRIDE; per_elem = oM + nq * 54;
        if (per_elem < 9 * nt * nt) per_elem = 9 * nt * nt;   // the staging area is reused for the local matrix (3 nt)^2
        per_elem += (per_elem & 1);
    }
};

template <int MC, int NTH = 256, int MINB = 1>
__global__ void __launch_bounds__(NTH, MINB) k_tangent_hypel_sym(const AsmParams p, int ntiles) {
    extern __shared__ __align__(16) double smem[];
    __shared__ unsigned char sTileN[160], sTileM0[160];
    const int tid = threadIdx.x, nth = blockDim.x;
    const int nt = p.nt, nq = p.nq, npe = p.npe;
    const HypelSymLayout L(npe, nq, nt);
    // tiles of one element: for every trial node N the test nodes 0..N in chunks of MC
    if (tid == 0) {
        int t = 0;
        for (int N = 0; N < nt; N++)
            for (int M0 = 0; M0 <= N; M0 += MC) { sTileN[t] = (unsigned char)N; sTileM0[t] = (unsigned char)M0; t++; }
    }
    for (int64_t base = (int64_t)blockIdx.x * p.EB; base < p.n_elems; base += (int64_t)gridDim.x * p.EB) {
        const int nb = (int)min((int64_t)p.EB, p.n_elems - base);
        __syncthreads();   // previous batch done with the staging area
        // coordinates
        for (int t = tid; t < nb * npe * 3; t += nth) {
            const int eb = t / (npe * 3), r = t % (npe * 3);
            smem[(size_t)eb * L.per_elem + L.oX + r] = p.coords[(size_t)p.conn[p.eid(base + eb) * npe + r / 3] * 3 + r % 3];
        }
        // current nodal values of the trial field (all threads; the displacement gradient below reads them from here)
        for (int t = tid; t < nb * nt * 3; t += nth) {
            const int eb = t / (nt * 3), r = t % (nt * 3);
            smem[(size_t)eb * L.per_elem + L.oU + r] = p.val_c[(size_t)p.ed_c[p.eid(base + eb) * nt + r / 3] * 3 + r % 3];
        }
        __syncthreads();
        // J^-T and det J per point (base/geometry.hpp:142-177,419-445)
        for (int t = tid; t < nb * nq; t += nth) {
            const int eb = t / nq, q = t % nq;
            double* E = smem + (size_t)eb * L.per_elem;
            const double* dN = p.dNg + (size_t)q * npe * 3;
            double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            for (int n = 0; n < npe; n++)
                for (int i = 0; i < 3; i++)
                    for (int a = 0; a < 3; a++) J[i][a] += E[L.oX + n * 3 + i] * dN[n * 3 + a];
            double aux[3][3], inv[3][3];
            for (int i = 0; i < 3; i++) for (int a = 0; a < 3; a++) aux[i][a] = J[a][i];
            E[L.oDet + q] = inv3(aux, inv);
            for (int i = 0; i < 3; i++) for (int a = 0; a < 3; a++) E[L.oCon + q * 9 + i * 3 + a] = inv[i][a];
        }
        __syncthreads();
        // physical gradients g_a = J^-T grad_xi phi_a
        for (int t = tid; t < nb * nq * nt; t += nth) {
            const int eb = t / (nq * nt), r = t % (nq * nt), q = r / nt;
            double* E = smem + (size_t)eb * L.per_elem;
            const double* con = E + L.oCon + q * 9;
            const double* dN = p.dNt + (size_t)r * 3;
#pragma unroll
            for (int c = 0; c < 3; c++) E[L.oG + r * 3 + c] = con[c * 3] * dN[0] + con[c * 3 + 1] * dN[1] + con[c * 3 + 2] * dN[2];
        }
        __syncthreads();
        // displacement gradient GradU(J,i) = sum_f g_f[J] u_f[i] per point, one thread per (point, J, i); summation over
        // the nodes in ascending order as in post/evaluateField.hpp:228-274 (parked in the F slot of the point)
        for (int t = tid; t < nb * nq * 9; t += nth) {
            const int eb = t / (nq * 9), r = t % (nq * 9), q = r / 9, J = (r % 9) / 3, i = r % 3;
            double* E = smem + (size_t)eb * L.per_elem;
            const double* g = E + L.oG + (size_t)q * nt * 3;
            const double* u = E + L.oU;
            double a = 0.;
            for (int f = 0; f < nt; f++) a += g[f * 3 + J] * u[f * 3 + i];
            E[L.oM + q * 54 + J * 3 + i] = a;
        }
        __syncthreads();
        // F, S, C per point (solid/Deformation.hpp:25-46, mat/hypel/*.hpp)
        for (int t = tid; t < nb * nq; t += nth) {
            const int eb = t / nq, q = t % nq;
            double* E = smem + (size_t)eb * L.per_elem;
            double GradU[3][3];
            for (int J = 0; J < 3; J++) for (int i = 0; i < 3; i++) GradU[J][i] = E[L.oM + q * 54 + J * 3 + i];
            double F[3][3], S[3][3], C[6][6];
            for (int i = 0; i < 3; i++) for (int J = 0; J < 3; J++) F[i][J] = (i == J ? 1. : 0.) + GradU[J][i];
            material_eval(p.kernel_id, p.p0, p.p1, F, S, C, true);
            double* m = E + L.oM + q * 54;
            for (int i = 0; i < 3; i++) for (int J = 0; J < 3; J++) { m[i * 3 + J] = F[i][J]; m[9 + i * 3 + J] = S[i][J]; }
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) m[18 + a * 6 + b] = C[a][b];
        }
        __syncthreads();
        // effective elasticity, one thread per (point, J, k, L)
        for (int t = tid; t < nb * nq * 27; t += nth) {
            const int eb = t / (nq * 27), r = t % (nq * 27), q = r / 27, jkl = r % 27, J = jkl / 9, k = (jkl / 3) % 3, Lx = jkl % 3;
            double* E = smem + (size_t)eb * L.per_elem;
            const double* m = E + L.oM + q * 54;
            double D[3];
#pragma unroll
            for (int A = 0; A < 3; A++) {
                double d = 0.;
#pragma unroll
                for (int B = 0; B < 3; B++) d = fma(m[18 + voigt_idx(A, J) * 6 + voigt_idx(B, Lx)], m[k * 3 + B], d);
                D[A] = d;
            }
            double* ce = E + L.oQ + q * HS_QSTRIDE;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                double v = (i == k) ? m[9 + J * 3 + Lx] : 0.;
#pragma unroll
                for (int A = 0; A < 3; A++) v = fma(m[i * 3 + A], D[A], v);
                ce[((i * 3 + J) * 3 + k) * 3 + Lx] = v;
            }
        }
        __syncthreads();
        // register tiles: exactly one per thread (the launch makes EB * ntiles <= blockDim.x)
        const int items = nb * ntiles;
        const bool has = tid < items;
        const int teb = has ? tid / ntiles : 0, tl = has ? tid % ntiles : 0;
        const int N = sTileN[tl], M0 = sTileM0[tl];
        double acc[MC][9];
#pragma unroll
        for (int m = 0; m < MC; m++)
#pragma unroll
            for (int x = 0; x < 9; x++) acc[m][x] = 0.;
        if (has) {
            const double* E = smem + (size_t)teb * L.per_elem;
            const double* G = E + L.oG;
            for (int q = 0; q < nq; q++) {
                const double wd = E[L.oDet + q] * p.w[q];
                const double* gN = G + ((size_t)q * nt + N) * 3;
                const double g0 = gN[0] * wd, g1 = gN[1] * wd, g2 = gN[2] * wd;
                const double* c = E + L.oQ + q * HS_QSTRIDE;   // the same address in all lanes working on this element: broadcast
                double T[27];   // [i][J][k]
#pragma unroll
                for (int x = 0; x < 27; x++) T[x] = fma(c[x * 3 + 2], g2, fma(c[x * 3 + 1], g1, c[x * 3] * g0));
#pragma unroll
                for (int m = 0; m < MC; m++) {
                    const int M = min(M0 + m, N);
                    const double* gM = G + ((size_t)q * nt + M) * 3;
                    const double h0 = gM[0], h1 = gM[1], h2 = gM[2];
#pragma unroll
                    for (int i = 0; i < 3; i++)
#pragma unroll
                        for (int k = 0; k < 3; k++)
                            acc[m][i * 3 + k] = fma(h2, T[(i * 3 + 2) * 3 + k], fma(h1, T[(i * 3 + 1) * 3 + k], fma(h0, T[(i * 3 + 0) * 3 + k], acc[m][i * 3 + k])));
                }
            }
        }
        __syncthreads();   // every thread is done with the staged gradients: the local matrix reuses that shared memory
        // local matrices -> shared memory (block and transposed block), then ONE coalesced pass over all entries:
        // consecutive threads read consecutive slots and add to neighbouring CSR entries
        const int nr = nt * 3, nn = nr * nr;
        if (has) {
            double* Kl = smem + (size_t)teb * L.per_elem;
#pragma unroll
            for (int m = 0; m < MC; m++) {
                const int M = M0 + m;
                if (M <= N) {
#pragma unroll
                    for (int i = 0; i < 3; i++)
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const double v = acc[m][i * 3 + k];
                            Kl[(M * 3 + i) * nr + N * 3 + k] = v;
                            if (M != N) Kl[(N * 3 + k) * nr + M * 3 + i] = v;
                        }
                }
            }
        }
        __syncthreads();
        for (int eb = 0; eb < nb; eb++) {
            const int64_t e = p.eid(base + eb);
            const double* Kl = smem + (size_t)eb * L.per_elem;
            if (p.kout) {   // atomic-free path: the local matrix goes to memory, the rows are gathered by k_gather_rows
                double* out = p.kout + (size_t)e * nn;
                for (int t = tid; t < nn; t += nth) out[t] = Kl[t];
                continue;
            }
            if (p.bbase) {
                // node-block positions: work item = (row i of the local matrix, trial node N): three neighbouring CSR
                // entries at base(M, N) + ci * len(M); the 5.8 KB (Q2) table of an element replaces 26 KB of slots
                const int64_t* bb = p.bbase + (size_t)e * nt * nt;
                const int32_t* bl = p.blen + (size_t)e * nt;
                for (int w = tid; w < nr * nt; w += nth) {
                    const int i = w / nt, N = w - i * nt, M = i / 3, ci = i - M * 3;
                    const int64_t b = __ldg(bb + M * nt + N);
                    const double* kl = Kl + i * nr + N * 3;
                    if (b >= 0) {
                        double* dst = p.val + b + (int64_t)ci * __ldg(bl + M);
                        atomicAdd(dst, kl[0]); atomicAdd(dst + 1, kl[1]); atomicAdd(dst + 2, kl[2]);
                    } else {
                        for (int k = 0; k < 3; k++) scatter_entry(p, e, i, N * 3 + k, nr, nr, kl[k]);
                    }
                }
                continue;
            }
            if (p.slot64) {   // systems with 2^31 or more non-zeros: 64-bit positions
                const int64_t* sw = p.slot64 + (size_t)e * nn;
                for (int t0 = tid; t0 < nn; t0 += 2 * nth) {
                    const int t1 = t0 + nth;
                    const int64_t a = __ldg(sw + t0), b = (t1 < nn) ? __ldg(sw + t1) : 0;
                    if (a >= 0) atomicAdd(p.val + a, Kl[t0]); else scatter_entry(p, e, t0 / nr, t0 % nr, nr, nr, Kl[t0]);
                    if (t1 < nn) { if (b >= 0) atomicAdd(p.val + b, Kl[t1]); else scatter_entry(p, e, t1 / nr, t1 % nr, nr, nr, Kl[t1]); }
                }
                continue;
            }
            const int32_t* sl = p.slot + (size_t)e * nn;
            constexpr int U = 4;   // slot loads in flight per thread
            for (int t0 = tid; t0 < nn; t0 += U * nth) {
                int32_t s4[U];
#pragma unroll
                for (int u = 0; u < U; u++) { const int t = t0 + u * nth; s4[u] = (t < nn) ? __ldg(sl + t) : 0; }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int t = t0 + u * nth;
                    if (t >= nn) continue;
                    if (s4[u] >= 0) atomicAdd(p.val + s4[u], Kl[t]);
                    else scatter_entry(p, e, t / nr, t % nr, nr, nr, Kl[t]);   // constrained row / column: lift or nothing
                }
            }
        }
    }
}
#endif
