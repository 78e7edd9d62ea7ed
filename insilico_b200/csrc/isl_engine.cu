// =============================================================================
// isl_engine.cu -- B200 (sm_100a) element-assembly engine behind the C ABI of
// include/insilico_b200.h.
//
// Data layout in HBM (all flat, SoA):
//   coords  f64 [n_nodes][dim]        conn     i32 [n_elems][npe]
//   per field: elem_dof i32 [n_elems][ndpe], eqn i32 [n_obj][ds] (-1 = not ACTIVE),
//              status u8, prescribed f64, values f64 [n_obj][ds]
//   system: rowptr i64 [n+1], col i32 [nnz] (ascending per row), val f64 [nnz], rhs f64 [n]
//   per (test,trial) pair: slot i32 [n_elems][nr][nc] = position of local entry (i,j) in val
//              (-1 when the row or the column is not ACTIVE)
//
// Kernels (FP64 CUDA cores; the per-element products are tiny, tensor cores do not apply):
//   k_tangent<DIM>   generic cooperative element kernel: coordinates, Jacobians and physical shape-function
//                    gradients staged in shared memory per element batch, then one thread per local
//                    matrix entry reduces over quadrature points and scatter-adds through the slot map
//                    (RED.ADD.F64); CONSTRAINED columns are lifted into rhs.
//   k_force<DIM>     residual forces / body force with the same staging.
//   k_q1hex_laplace  specialised hot path for Q1 hex scalar Laplace (BASELINE config 2), one thread per
//                    element with the 8x8 symmetric local matrix in registers.
//   pattern build    (row,col) 64-bit keys -> cub radix sort -> unique -> CSR; slot map by binary search.
// =============================================================================
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/insilico_b200.h"
#include "isl_dof.hpp"
#include "isl_tables.hpp"
#include "isl_constraints.hpp"

namespace {

thread_local std::string g_error;

struct IslError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define ISL_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            throw IslError(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + \
                           std::to_string(__LINE__));                                                    \
    } while (0)

#define ISL_REQUIRE(cond, msg) \
    do { if (!(cond)) throw IslError(std::string(msg)); } while (0)

template <class F>
int guarded(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { g_error = e.what(); return 1; }
    catch (...) { g_error = "unknown error"; return 1; }
}

// simple owning device buffer
template <class T>
struct DevBuf {
    T* p = nullptr; size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    void alloc(size_t count) {
        if (count == n && p) return;
        release();
        if (count) ISL_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
        n = count;
    }
    void swap(DevBuf& o) { std::swap(p, o.p); std::swap(n, o.n); }
};

// ---------------------------------------------------------------------------------------------
// device-side parameter blocks
struct AsmParams {
    // mesh
    const double* coords; const int32_t* conn; int64_t n_elems; int npe;
    // tables
    const double* w; const double* dNg; const double* Ng;
    const double* Nt; const double* dNt; const double* Nc; const double* dNc;
    int nq, nt, nc, dst, dsc, bubnov;
    // fields
    const int32_t* ed_t; const int32_t* ed_c;
    const int32_t* eqn_t; const int32_t* eqn_c;
    const uint8_t* st_c; const double* presc_c; const double* val_c;
    // system
    const int32_t* slot; const int64_t* slot64;   // element -> CSR position maps: 32-bit, or 64-bit when nnz >= 2^31 (exactly one is set)
    // node-block position map of the hyperelastic tile kernel (one base position per pair of nodes instead of ds*ds slots)
    const int64_t* bbase; const int32_t* blen;
    double* kout;   // hyperelastic tile kernel: local matrices go to kout[e][nr][nr] instead of the CSR (isl_gather.cuh);
                    // k_tangent: kout[e][nt][nc] (one scalar per node pair) or kout[e][nt dst][nc dsc] (Stokes coupling blocks)
    double* val; double* rhs;
    int kernel_id; double p0, p1; int incremental; double factor;
    int EB; int need_gt, need_gc, nqdata;
    int tile;   // k_tangent: register strips (one thread = one test function x up to five trial functions, or all components of a coupling entry)
    int body; double f[3];
    const double* fq;   // body == 2: force sampled at the quadrature points, [n_elems][nq][dst]
    const double* kq;   // Laplace kernels: conductivity sampled at the quadrature points, [n_elems][nq] (nullptr: constant p0)
    // third field of the tuple (AuxField1; fluid::Convection: the advection velocity), same FE basis as the trial field
    const int32_t* ed_a; const double* val_a; int dsa;
    // linear constraints with master DoFs (nullptr when the field has none): per DoF component k the masters
    // [cptr[k], cptr[k+1]) as equation numbers cm[] with weights cw[]; CSR pattern for the entries they reach
    const int32_t* cptr_t; const int32_t* cm_t; const double* cw_t;
    const int32_t* cptr_c; const int32_t* cm_c; const double* cw_c;
    const int64_t* rowptr; const int32_t* col;
    // elements are processed in a locality-sorted order (smallest equation number of the element): neighbouring
    // batches scatter into neighbouring CSR rows whatever the caller's element order is (nullptr: as given)
    const int32_t* eorder;
    __device__ __forceinline__ int64_t eid(int64_t k) const { return eorder ? (int64_t)eorder[k] : k; }
};

__device__ __forceinline__ int voigt_idx(int i, int j) {
    // mat/TensorAlgebra.hpp:149-164
    const int map[9] = {0, 3, 4, 1, -1, 5, -1, -1, 2};
    return map[(i + 1) * (j + 1) - 1];
}

__device__ __forceinline__ double inv3(const double m[3][3], double inv[3][3]) {
    const double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
    const double c10 = m[2][1] * m[0][2] - m[2][2] * m[0][1];
    const double c20 = m[0][1] * m[1][2] - m[0][2] * m[1][1];
    const double det = c00 * m[0][0] + (c10 * m[1][0] + c20 * m[2][0]);
    const double id = 1.0 / det;
    inv[0][0] = c00 * id; inv[0][1] = c10 * id; inv[0][2] = c20 * id;
    inv[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) * id;
    inv[1][1] = (m[2][2] * m[0][0] - m[2][0] * m[0][2]) * id;
    inv[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
    inv[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) * id;
    inv[2][1] = (m[2][0] * m[0][1] - m[2][1] * m[0][0]) * id;
    inv[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
    return det;
}

// material response at a quadrature point: S (2nd PK) and C (6x6 Voigt)
// mat/hypel/StVenant.hpp:77-114, mat/hypel/NeoHookeanCompressible.hpp:77-172
__device__ void material_eval(int kid, double lambda, double mu, const double F[3][3], double S[3][3], double C[6][6],
                              bool want_C) {
    double CG[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) CG[i][j] = F[0][i] * F[0][j] + F[1][i] * F[1][j] + F[2][i] * F[2][j];
    if (kid == ISL_K_HYPEL_STVENANT) {
        double E[3][3];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) E[i][j] = 0.5 * (CG[i][j] - (i == j ? 1. : 0.));
        const double trE = E[0][0] + E[1][1] + E[2][2];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) S[i][j] = lambda * trE * (i == j ? 1. : 0.) + 2. * mu * E[i][j];
        if (want_C) {
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) C[i][j] = 0.;
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[i][j] = lambda;
            for (int i = 0; i < 3; i++) C[i][i] += 2. * mu;
            for (int i = 3; i < 6; i++) C[i][i] += mu;
        }
    } else {
        double Ci[3][3];
        inv3(CG, Ci);
        const double J = F[0][0] * F[1][1] * F[2][2] + F[0][1] * F[1][2] * F[2][0] + F[0][2] * F[1][0] * F[2][1] -
                         F[0][0] * F[1][2] * F[2][1] - F[0][1] * F[1][0] * F[2][2] - F[0][2] * F[1][1] * F[2][0];
        const double logJ = log(J);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) S[i][j] = (lambda * logJ - mu) * Ci[i][j] + mu * (i == j ? 1. : 0.);
        if (want_C) {
            const double fac2 = mu - lambda * logJ;
            for (int A = 0; A < 3; A++)
                for (int B = A; B < 3; B++)
                    for (int Cc = 0; Cc < 3; Cc++)
                        for (int D = Cc; D < 3; D++)
                            C[voigt_idx(A, B)][voigt_idx(Cc, D)] =
                                lambda * Ci[A][B] * Ci[Cc][D] + fac2 * (Ci[A][Cc] * Ci[B][D] + Ci[A][D] * Ci[B][Cc]);
        }
    }
}

// shared-memory staging common to tangent and force kernels.  Layout (doubles):
//   sX [EB][npe][DIM] | sCon [EB][nq][DIM*DIM] | sDet [EB][nq] | sGt [EB][nq][nt][DIM] | sGc [...] | sQ [EB][nq][nqdata]
template <int DIM>
struct Stage {
    double *sX, *sCon, *sDet, *sGt, *sGc, *sQ;
    __device__ Stage(double* base, const AsmParams& p) {
        sX = base; base += (size_t)p.EB * p.npe * DIM;
        sCon = base; base += (size_t)p.EB * p.nq * DIM * DIM;
        sDet = base; base += (size_t)p.EB * p.nq;
        sGt = base; if (p.need_gt) base += (size_t)p.EB * p.nq * p.nt * DIM;
        if (p.need_gc && !(p.bubnov && p.need_gt)) { sGc = base; base += (size_t)p.EB * p.nq * p.nc * DIM; }
        else sGc = sGt;
        sQ = base;
    }
};

// geometry + gradients of one element batch:
//   J(i,a) = sum_n x_n[i] dphi_n/dxi_a ; contra = J^{-T} ; grad_x phi = contra * grad_xi phi
template <int DIM>
__device__ void stage_batch(const AsmParams& p, const Stage<DIM>& s, int64_t base, int nb) {
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int t = tid; t < nb * p.npe * DIM; t += nth) {
        const int eb = t / (p.npe * DIM), r = t % (p.npe * DIM);
        const int a = r / DIM, d = r % DIM;
        s.sX[t] = p.coords[(size_t)p.conn[(p.eid(base + eb)) * p.npe + a] * DIM + d];
    }
    __syncthreads();
    for (int t = tid; t < nb * p.nq; t += nth) {
        const int eb = t / p.nq, q = t % p.nq;
        const double* X = s.sX + (size_t)eb * p.npe * DIM;
        const double* dN = p.dNg + (size_t)q * p.npe * DIM;
        double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int n = 0; n < p.npe; n++)
            for (int i = 0; i < DIM; i++)
                for (int a = 0; a < DIM; a++) J[i][a] += X[n * DIM + i] * dN[n * DIM + a];
        double* con = s.sCon + (size_t)t * DIM * DIM;
        double det;
        if (DIM == 3) {
            double aux[3][3], inv[3][3];
            for (int i = 0; i < 3; i++) for (int a = 0; a < 3; a++) aux[i][a] = J[a][i];
            det = inv3(aux, inv);
            for (int i = 0; i < 3; i++) for (int a = 0; a < 3; a++) con[i * 3 + a] = inv[i][a];
        } else {
            // aux = J^T ; 2x2 inverse
            const double a00 = J[0][0], a01 = J[1][0], a10 = J[0][1], a11 = J[1][1];
            det = a00 * a11 - a10 * a01;
            const double id = 1.0 / det;
            con[0] = a11 * id; con[2] = -a10 * id; con[1] = -a01 * id; con[3] = a00 * id;
        }
        s.sDet[t] = det;
    }
    __syncthreads();
    if (p.need_gt)
        for (int t = tid; t < nb * p.nq * p.nt; t += nth) {
            const int eq = t / p.nt, a = t % p.nt, q = eq % p.nq;
            const double* con = s.sCon + (size_t)eq * DIM * DIM;
            const double* dN = p.dNt + ((size_t)q * p.nt + a) * DIM;
            for (int r = 0; r < DIM; r++) {
                double v = con[r * DIM] * dN[0];
                for (int c = 1; c < DIM; c++) v += con[r * DIM + c] * dN[c];
                s.sGt[(size_t)t * DIM + r] = v;
            }
        }
    if (p.need_gc && !(p.bubnov && p.need_gt))
        for (int t = tid; t < nb * p.nq * p.nc; t += nth) {
            const int eq = t / p.nc, a = t % p.nc, q = eq % p.nq;
            const double* con = s.sCon + (size_t)eq * DIM * DIM;
            const double* dN = p.dNc + ((size_t)q * p.nc + a) * DIM;
            for (int r = 0; r < DIM; r++) {
                double v = con[r * DIM] * dN[0];
                for (int c = 1; c < DIM; c++) v += con[r * DIM + c] * dN[c];
                s.sGc[(size_t)t * DIM + r] = v;
            }
        }
    __syncthreads();
}

// displacement / velocity gradient of the trial field at (eb,q): GradU(J,i) = sum_f g_f[J] u_f[i]
template <int DIM>
__device__ void trial_gradient(const AsmParams& p, const Stage<DIM>& s, int64_t e, int eq, double GradU[3][3]) {
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) GradU[a][b] = 0.;
    const double* g = s.sGc + (size_t)eq * p.nc * DIM;
    for (int f = 0; f < p.nc; f++) {
        const int32_t obj = p.ed_c[e * p.nc + f];
        for (int J = 0; J < DIM; J++)
            for (int i = 0; i < p.dsc; i++) GradU[J][i] += g[f * DIM + J] * p.val_c[(size_t)obj * p.dsc + i];
    }
}

// the same for the auxiliary field (same basis as the trial field), and field values u(xi_q) = sum_f phi_f u_f
// (base/post/evaluateField.hpp)
template <int DIM>
__device__ void aux_gradient(const AsmParams& p, const Stage<DIM>& s, int64_t e, int eq, double GradU[3][3]) {
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) GradU[a][b] = 0.;
    const double* g = s.sGc + (size_t)eq * p.nc * DIM;
    for (int f = 0; f < p.nc; f++) {
        const int32_t obj = p.ed_a[e * p.nc + f];
        for (int J = 0; J < DIM; J++)
            for (int i = 0; i < p.dsa; i++) GradU[J][i] += g[f * DIM + J] * p.val_a[(size_t)obj * p.dsa + i];
    }
}
__device__ __forceinline__ void field_value(const AsmParams& p, const int32_t* ed, const double* val, int ds, int64_t e, int q, double u[3]) {
    for (int i = 0; i < 3; i++) u[i] = 0.;
    for (int f = 0; f < p.nc; f++) {
        const int32_t obj = ed[e * p.nc + f];
        for (int i = 0; i < ds; i++) u[i] += p.Nc[q * p.nc + f] * val[(size_t)obj * ds + i];
    }
}

template <int DIM>
__device__ void deformation_gradient(const AsmParams& p, const Stage<DIM>& s, int64_t e, int eq, double F[3][3]) {
    double GradU[3][3];
    trial_gradient<DIM>(p, s, e, eq, GradU);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F[i][j] = (i == j ? 1. : 0.);
    for (int i = 0; i < p.dsc; i++) for (int J = 0; J < DIM; J++) F[i][J] += GradU[J][i];
}

__device__ __forceinline__ int64_t find_in_row(const int64_t* rowptr, const int32_t* col, int32_t r, int32_t c) {
    return isl_find_in_row(rowptr, col, r, c);
}
struct DeviceAdd {
    __device__ static void add(double* target, double value) { atomicAdd(target, value); }
};

// entry (kr, k) of an element whose row and/or column DoF is not ACTIVE while the field has slaves of master DoFs:
// isl_constraints.hpp (positions are looked up in the CSR row; the slot map only holds ACTIVE x ACTIVE); rare path
__device__ __noinline__ void scatter_constrained(const AsmParams& p, size_t kr, int32_t r, size_t k, double v) {
    const bool c_con = (p.st_c[k] == ISL_CONSTRAINED);
    const double g = c_con ? (p.incremental ? p.presc_c[k] - p.val_c[k] : p.presc_c[k]) : 0.;
    const IslMasters mt{p.cptr_t, p.cm_t, p.cw_t}, mc{p.cptr_c, p.cm_c, p.cw_c};
    isl_scatter_constrained<DeviceAdd>(mt, mc, p.rowptr, p.col, p.val, p.rhs, kr, r, k, p.eqn_c[k], c_con, g, v);
}

// scatter one local matrix entry: ACTIVE x ACTIVE -> CSR value,
// ACTIVE row x CONSTRAINED column -> rhs -= g * K; slaves of master DoFs -> scatter_constrained
__device__ __forceinline__ void scatter_entry(const AsmParams& p, int64_t e, int i, int j, int nr, int ncl, double v) {
    const int M = i / p.dst, ci = i % p.dst;
    const int N = j / p.dsc, cj = j % p.dsc;
    if (p.slot64 || p.slot) {
        const size_t idx = ((size_t)e * nr + i) * ncl + j;
        const int64_t sl = p.slot64 ? p.slot64[idx] : (int64_t)p.slot[idx];
        if (sl >= 0) { atomicAdd(p.val + sl, v); return; }
    }
    const size_t kr = (size_t)p.ed_t[e * p.nt + M] * p.dst + ci;
    const int32_t r = p.eqn_t[kr];
    const size_t k = (size_t)p.ed_c[e * p.nc + N] * p.dsc + cj;
    if (!p.slot64 && !p.slot && r >= 0 && p.eqn_c[k] >= 0 && p.cptr_t == nullptr && p.cptr_c == nullptr) {
        // no slot map (node-block map of the tile kernel, irregular pair): the position is searched in the row
        const int64_t pos = find_in_row(p.rowptr, p.col, r, p.eqn_c[k]);
        if (pos >= 0) atomicAdd(p.val + pos, v);
        return;
    }
    if (p.cptr_t != nullptr || p.cptr_c != nullptr) { scatter_constrained(p, kr, r, k, v); return; }
    if (r < 0) return;
    if (p.st_c[k] == ISL_CONSTRAINED) {
        const double g = p.incremental ? p.presc_c[k] - p.val_c[k] : p.presc_c[k];
        atomicAdd(p.rhs + r, -(g * v));
    }
}

// entry (M,i; N,k) through the node-block map when the pair is regular, else the per-entry path
__device__ __forceinline__ void scatter_block_entry(const AsmParams& p, int64_t e, int M, int i, int N, int k, int nr, int ncl, double v) {
    if (p.bbase) {
        const int64_t b = p.bbase[((size_t)e * p.nt + M) * p.nc + N];
        if (b >= 0) { atomicAdd(p.val + b + (int64_t)i * p.blen[(size_t)e * p.nt + M] + k, v); return; }
    }
    scatter_entry(p, e, M * p.dst + i, N * p.dsc + k, nr, ncl, v);
}

// entry K[M][N] of an integrand that repeats one scalar on every DoF component (Laplace, Mass, Convection)
__device__ __forceinline__ void emit_node_pair(const AsmParams& p, int64_t e, int M, int N, int nr, int ncl, double v) {
    if (p.kout) { p.kout[((size_t)e * p.nt + M) * p.nc + N] = v; return; }   // atomic-free path: rows gathered by k_gen_gather_rows
    for (int c = 0; c < p.dsc; c++) scatter_block_entry(p, e, M, c, N, c, nr, ncl, v);
}
// entry (M,i; N,k) of a coupling block
__device__ __forceinline__ void emit_entry(const AsmParams& p, int64_t e, int M, int i, int N, int k, int nr, int ncl, double v) {
    if (p.kout) { p.kout[((size_t)e * nr + M * p.dst + i) * ncl + N * p.dsc + k] = v; return; }
    scatter_block_entry(p, e, M, i, N, k, nr, ncl, v);
}

template <int DIM>
__global__ void __launch_bounds__(256) k_tangent(const AsmParams p) {
    extern __shared__ double smem[];
    const Stage<DIM> s(smem, p);
    const int tid = threadIdx.x, nth = blockDim.x;
    const int nr = p.nt * p.dst, ncl = p.nc * p.dsc;
    for (int64_t base = (int64_t)blockIdx.x * p.EB; base < p.n_elems; base += (int64_t)gridDim.x * p.EB) {
        const int nb = (int)min((int64_t)p.EB, p.n_elems - base);
        stage_batch<DIM>(p, s, base, nb);
        if (p.kernel_id == ISL_K_HYPEL_STVENANT || p.kernel_id == ISL_K_HYPEL_NEOHOOKE) {
            // effective elasticity per quadrature point, hoisted out of the (M,N) loops
            // (solid/HyperElastic.hpp:282-310 evaluates it per entry)
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
            for (int t = tid; t < nb * nr * ncl; t += nth) {
                const int eb = t / (nr * ncl), ij = t % (nr * ncl), i = ij / ncl, j = ij % ncl;
                const int M = i / DIM, ci = i % DIM, N = j / DIM, ck = j % DIM;
                double acc = 0.;
                for (int q = 0; q < p.nq; q++) {
                    const int eq = eb * p.nq + q;
                    const double* gM = s.sGt + ((size_t)eq * p.nt + M) * DIM;
                    const double* gN = s.sGc + ((size_t)eq * p.nc + N) * DIM;
                    const double* ce = s.sQ + (size_t)eq * 81 + (ci * 3) * 9 + ck * 3;
                    double sum = 0.;
                    for (int J = 0; J < DIM; J++)
                        for (int L = 0; L < DIM; L++) sum += gM[J] * ce[J * 9 + L] * gN[L];
                    acc += sum * (s.sDet[eq] * p.w[q]);
                }
                scatter_block_entry(p, p.eid(base + eb), M, ci, N, ck, nr, ncl, acc);
            }
        } else if (p.kernel_id == ISL_K_LAPLACE || p.kernel_id == ISL_K_VECTOR_LAPLACE) {
            for (int t = tid; t < nb * p.nt * p.nc; t += nth) {
                const int eb = t / (p.nt * p.nc), mn = t % (p.nt * p.nc), M = mn / p.nc, N = mn % p.nc;
                double acc = 0.;
                for (int q = 0; q < p.nq; q++) {
                    const int eq = eb * p.nq + q;
                    const double* gM = s.sGt + ((size_t)eq * p.nt + M) * DIM;
                    const double* gN = s.sGc + ((size_t)eq * p.nc + N) * DIM;
                    double dot = gM[0] * gN[0];
                    for (int k = 1; k < DIM; k++) dot += gM[k] * gN[k];
                    const double kap = p.kq ? p.kq[(size_t)p.eid(base + eb) * p.nq + q] : p.p0;
                    acc += dot * (kap * s.sDet[eq] * p.w[q]);
                }
                emit_node_pair(p, p.eid(base + eb), M, N, nr, ncl, acc);
            }
        } else if (p.kernel_id == ISL_K_MASS) {
            // base/kernel/Mass.hpp:88-138: (factor detJ w) phi_M psi_N on every DoF component
            for (int t = tid; t < nb * p.nt * p.nc; t += nth) {
                const int eb = t / (p.nt * p.nc), mn = t % (p.nt * p.nc), M = mn / p.nc, N = mn % p.nc;
                double acc = 0.;
                for (int q = 0; q < p.nq; q++)
                    acc += (p.p0 * s.sDet[eb * p.nq + q] * p.w[q]) * p.Nt[q * p.nt + M] * p.Nc[q * p.nc + N];
                emit_node_pair(p, p.eid(base + eb), M, N, nr, ncl, acc);
            }
        } else if (p.kernel_id == ISL_K_CONVECTION) {
            // fluid/Convection.hpp:88-166 (Picard form): phi_M (uAdv . grad phi_N + 0.5 div(u) phi_N) rho detJ w on every
            // component; uAdv from the auxiliary field, div(u) of the trial field's current state, hoisted per point
            for (int t = tid; t < nb * p.nq; t += nth) {
                const int eb = t / p.nq, q = t % p.nq;
                const int64_t e = p.eid(base + eb);
                double u[3], G[3][3];
                field_value(p, p.ed_a, p.val_a, p.dsa, e, q, u);
                trial_gradient<DIM>(p, s, e, t, G);
                double div = 0.;
                for (int d = 0; d < p.dsc; d++) div += G[d][d];
                double* qd = s.sQ + (size_t)t * 4;
                qd[0] = u[0]; qd[1] = u[1]; qd[2] = u[2]; qd[3] = div;
            }
            __syncthreads();
            for (int t = tid; t < nb * p.nt * p.nc; t += nth) {
                const int eb = t / (p.nt * p.nc), mn = t % (p.nt * p.nc), M = mn / p.nc, N = mn % p.nc;
                double acc = 0.;
                for (int q = 0; q < p.nq; q++) {
                    const int eq = eb * p.nq + q;
                    const double* gN = s.sGc + ((size_t)eq * p.nc + N) * DIM;
                    const double* qd = s.sQ + (size_t)eq * 4;
                    double adv = 0.;
                    for (int k = 0; k < p.dst; k++) adv += qd[k] * gN[k];
                    acc += p.Nt[q * p.nt + M] * (adv + 0.5 * qd[3] * p.Nc[q * p.nc + N]) * p.p0 * s.sDet[eq] * p.w[q];
                }
                emit_node_pair(p, p.eid(base + eb), M, N, nr, ncl, acc);
            }
        } else if (p.kernel_id == ISL_K_PRESSURE_GRADIENT) {
            // B(M d + i, N) = -detJ w g_M[i] psi_N
            for (int t = tid; t < nb * nr * ncl; t += nth) {
                const int eb = t / (nr * ncl), ij = t % (nr * ncl), i = ij / ncl, N = ij % ncl;
                const int M = i / p.dst, d = i % p.dst;
                double acc = 0.;
                if (d < DIM)
                    for (int q = 0; q < p.nq; q++) {
                        const int eq = eb * p.nq + q;
                        acc += -s.sDet[eq] * p.w[q] * s.sGt[((size_t)eq * p.nt + M) * DIM + d] * p.Nc[q * p.nc + N];
                    }
                emit_entry(p, p.eid(base + eb), M, d, N, 0, nr, ncl, acc);
            }
        } else if (p.kernel_id == ISL_K_VELOCITY_DIVERGENCE) {
            // transpose of the pressure-gradient block on the transposed tuple, optional sign change
            const double sgn = (p.p0 != 0.) ? -1.0 : 1.0;
            for (int t = tid; t < nb * nr * ncl; t += nth) {
                const int eb = t / (nr * ncl), ij = t % (nr * ncl), Mp = ij / ncl, j = ij % ncl;
                const int N = j / p.dsc, d = j % p.dsc;
                double acc = 0.;
                if (d < DIM)
                    for (int q = 0; q < p.nq; q++) {
                        const int eq = eb * p.nq + q;
                        acc += -s.sDet[eq] * p.w[q] * s.sGc[((size_t)eq * p.nc + N) * DIM + d] * p.Nt[q * p.nt + Mp];
                    }
                emit_entry(p, p.eid(base + eb), Mp, 0, N, d, nr, ncl, sgn * acc);
            }
        }
        __syncthreads();
    }
}

// k_tangent with register strips for the Laplace-type integrands and the Stokes coupling blocks (knob gen_tile): the
// per-entry loops of k_tangent spend half of their instructions on index arithmetic and on loads the neighbouring
// entries repeat (profiles/r2/r2_t_ncu_C5_gather.md); same staging, same sum order per entry
template <int DIM>
__global__ void __launch_bounds__(256, 3) k_tangent_strips(const AsmParams p) {
    extern __shared__ double smem[];
    const Stage<DIM> s(smem, p);
    const int tid = threadIdx.x, nth = blockDim.x;
    const int nr = p.nt * p.dst, ncl = p.nc * p.dsc;
    for (int64_t base = (int64_t)blockIdx.x * p.EB; base < p.n_elems; base += (int64_t)gridDim.x * p.EB) {
        const int nb = (int)min((int64_t)p.EB, p.n_elems - base);
        stage_batch<DIM>(p, s, base, nb);
        if (p.kernel_id == ISL_K_LAPLACE || p.kernel_id == ISL_K_VECTOR_LAPLACE) {
            // one thread = test function M x a strip of up to five trial functions: the gradient of M, the weight and the
            // index arithmetic are shared by the strip (same sum order per entry as the per-entry loop below)
            constexpr int TN = 5;
            const int strips = (p.nc + TN - 1) / TN, per = p.nt * strips;
            for (int t = tid; t < nb * per; t += nth) {
                const int eb = t / per, ms = t - eb * per, M = ms / strips, N0 = (ms - M * strips) * TN;
                const int cnt = min(TN, p.nc - N0);
                const int64_t e = p.eid(base + eb);
                double acc[TN];
#pragma unroll
                for (int j = 0; j < TN; j++) acc[j] = 0.;
                for (int q = 0; q < p.nq; q++) {
                    const int eq = eb * p.nq + q;
                    const double* gM = s.sGt + ((size_t)eq * p.nt + M) * DIM;
                    const double* gN = s.sGc + ((size_t)eq * p.nc + N0) * DIM;
                    const double kap = p.kq ? p.kq[(size_t)e * p.nq + q] : p.p0;
                    const double f = kap * s.sDet[eq] * p.w[q];
                    double m[DIM];
#pragma unroll
                    for (int k = 0; k < DIM; k++) m[k] = gM[k];
#pragma unroll
                    for (int j = 0; j < TN; j++)
                        if (j < cnt) {
                            double dot = m[0] * gN[j * DIM];
#pragma unroll
                            for (int k = 1; k < DIM; k++) dot += m[k] * gN[j * DIM + k];
                            acc[j] += dot * f;
                        }
                }
#pragma unroll
                for (int j = 0; j < TN; j++)
                    if (j < cnt) emit_node_pair(p, e, M, N0 + j, nr, ncl, acc[j]);
            }
        } else if (p.kernel_id == ISL_K_PRESSURE_GRADIENT) {
            // one thread = node pair (M, N), all components d of the test gradient
            const int per = p.nt * p.nc;
            for (int t = tid; t < nb * per; t += nth) {
                const int eb = t / per, mn = t - eb * per, M = mn / p.nc, N = mn - M * p.nc;
                double acc[DIM];
#pragma unroll
                for (int d = 0; d < DIM; d++) acc[d] = 0.;
                for (int q = 0; q < p.nq; q++) {
                    const int eq = eb * p.nq + q;
                    const double dw = -s.sDet[eq] * p.w[q], psi = p.Nc[q * p.nc + N];
                    const double* gM = s.sGt + ((size_t)eq * p.nt + M) * DIM;
#pragma unroll
                    for (int d = 0; d < DIM; d++) acc[d] += dw * gM[d] * psi;
                }
                const int64_t e = p.eid(base + eb);
#pragma unroll
                for (int d = 0; d < DIM; d++) emit_entry(p, e, M, d, N, 0, nr, ncl, acc[d]);
            }
        } else if (p.kernel_id == ISL_K_VELOCITY_DIVERGENCE) {
            const double sgn = (p.p0 != 0.) ? -1.0 : 1.0;
            const int per = p.nt * p.nc;
            for (int t = tid; t < nb * per; t += nth) {
                const int eb = t / per, mn = t - eb * per, Mp = mn / p.nc, N = mn - Mp * p.nc;
                double acc[DIM];
#pragma unroll
                for (int d = 0; d < DIM; d++) acc[d] = 0.;
                for (int q = 0; q < p.nq; q++) {
                    const int eq = eb * p.nq + q;
                    const double dw = -s.sDet[eq] * p.w[q], phi = p.Nt[q * p.nt + Mp];
                    const double* gN = s.sGc + ((size_t)eq * p.nc + N) * DIM;
#pragma unroll
                    for (int d = 0; d < DIM; d++) acc[d] += dw * gN[d] * phi;
                }
                const int64_t e = p.eid(base + eb);
#pragma unroll
                for (int d = 0; d < DIM; d++) emit_entry(p, e, Mp, 0, N, d, nr, ncl, sgn * acc[d]);
            }
        }
        __syncthreads();
    }
}

// residual forces and body force
template <int DIM>
__global__ void __launch_bounds__(256) k_force(const AsmParams p) {
    extern __shared__ double smem[];
    const Stage<DIM> s(smem, p);
    const int tid = threadIdx.x, nth = blockDim.x;
    const int nr = p.nt * p.dst;
    for (int64_t base = (int64_t)blockIdx.x * p.EB; base < p.n_elems; base += (int64_t)gridDim.x * p.EB) {
        const int nb = (int)min((int64_t)p.EB, p.n_elems - base);
        stage_batch<DIM>(p, s, base, nb);
        if (!p.body) {
            for (int t = tid; t < nb * p.nq; t += nth) {
                const int eb = t / p.nq, q = t % p.nq;
                const int64_t e = p.eid(base + eb);
                double* qd = s.sQ + (size_t)t * 9;
                if (p.kernel_id == ISL_K_HYPEL_STVENANT || p.kernel_id == ISL_K_HYPEL_NEOHOOKE) {
                    double F[3][3], S[3][3], C[6][6];
                    deformation_gradient<DIM>(p, s, e, t, F);
                    material_eval(p.kernel_id, p.p0, p.p1, F, S, C, false);
                    for (int i = 0; i < 3; i++)
                        for (int j = 0; j < 3; j++) qd[i * 3 + j] = F[i][0] * S[0][j] + F[i][1] * S[1][j] + F[i][2] * S[2][j];
                } else if (p.kernel_id == ISL_K_CONVECTION) {   // fluid/Convection.hpp:170-220: (U . grad) u_aux
                    double U[3], G[3][3];
                    field_value(p, p.ed_c, p.val_c, p.dsc, e, q, U);
                    aux_gradient<DIM>(p, s, e, t, G);
                    for (int i = 0; i < 3; i++) {
                        double cd = 0.;
                        for (int k = 0; k < p.dst; k++) cd += U[k] * G[k][i];
                        qd[i] = cd;
                    }
                } else if (p.kernel_id == ISL_K_PRESSURE_GRADIENT) {
                    double pr = 0.;
                    for (int f = 0; f < p.nc; f++) pr += p.Nc[q * p.nc + f] * p.val_c[(size_t)p.ed_c[e * p.nc + f] * p.dsc];
                    qd[0] = pr;
                } else {
                    double G[3][3];
                    trial_gradient<DIM>(p, s, e, t, G);
                    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) qd[a * 3 + b] = G[a][b];
                }
            }
            __syncthreads();
        }
        for (int t = tid; t < nb * nr; t += nth) {
            const int eb = t / nr, i = t % nr, M = i / p.dst, ci = i % p.dst;
            const int64_t e = p.eid(base + eb);
            double acc = 0.;
            for (int q = 0; q < p.nq; q++) {
                const int eq = eb * p.nq + q;
                const double detJ = s.sDet[eq], w = p.w[q];
                const double* qd = s.sQ + (size_t)eq * 9;
                if (p.body) {
                    const double fc = (p.body == 2) ? p.fq[((size_t)e * p.nq + q) * p.dst + ci] : p.f[ci];
                    acc += fc * p.Nt[q * p.nt + M] * w * detJ;
                    continue;
                }
                const double* gM = s.sGt + ((size_t)eq * p.nt + M) * DIM;
                switch (p.kernel_id) {
                    case ISL_K_LAPLACE: {  // heat/Laplace.hpp:153-181
                        double dot = (p.p0 * qd[0]) * gM[0];
                        for (int d = 1; d < DIM; d++) dot += (p.p0 * qd[d * 3]) * gM[d];
                        acc += dot * detJ * w;
                    } break;
                    case ISL_K_VECTOR_LAPLACE: {  // fluid/VectorLaplace.hpp:78-106
                        double dot = 0.;
                        for (int k = 0; k < DIM; k++) dot += qd[k * 3 + ci] * gM[k];
                        acc += p.p0 * dot * detJ * w;
                    } break;
                    case ISL_K_HYPEL_STVENANT:
                    case ISL_K_HYPEL_NEOHOOKE: {  // solid/HyperElastic.hpp:209-257
                        double sum = 0.;
                        for (int J = 0; J < DIM; J++) sum += qd[ci * 3 + J] * gM[J];
                        acc += sum * (detJ * w);
                    } break;
                    case ISL_K_PRESSURE_GRADIENT:  // fluid/PressureGradient.hpp:130-155
                        if (ci < DIM) acc += -gM[ci] * qd[0] * detJ * w;
                        break;
                    case ISL_K_CONVECTION:
                        acc += p.p0 * qd[ci] * p.Nt[q * p.nt + M] * detJ * w;
                        break;
                    case ISL_K_VELOCITY_DIVERGENCE: {  // fluid/VelocityDivergence.hpp:100-124
                        double div = 0.;
                        for (int d = 0; d < DIM; d++) div += qd[d * 3 + d];
                        acc += ((p.p0 != 0.) ? -1.0 : 1.0) * p.Nt[q * p.nt + M] * div * detJ * w;
                    } break;
                }
            }
            const size_t kr = (size_t)p.ed_t[e * p.nt + M] * p.dst + ci;
            const int32_t r = p.eqn_t[kr];
            if (r >= 0) atomicAdd(p.rhs + r, p.factor * acc);
            else if (p.cptr_t != nullptr)  // slave of master DoFs (asmb/assembleForces.hpp:118-131)
                isl_scatter_force_to_masters<DeviceAdd>(IslMasters{p.cptr_t, p.cm_t, p.cw_t}, p.rhs, kr, p.factor * acc);
        }
        __syncthreads();
    }
}

#include "isl_tangent_tiled.cuh"
#include "isl_gather.cuh"
#include "isl_neumann.cuh"
#include "isl_dof_dev.cuh"

// ---------------------------------------------------------------------------------------------
// specialised hot path: Q1 hex geometry, Q1 scalar field, Laplace, 2x2x2 Gauss rule (BASELINE config 2).
// One thread per element; symmetric 8x8 local matrix (36 accumulators) in registers; scatter through the
// slot map with RED.ADD.F64; Dirichlet lift fused.  Tables live in constant memory.
__constant__ double c_q1_dN[8 * 8 * 3];  // [q][a][d] reference gradients at the 8 Gauss points
__constant__ double c_q1_w[8];
__constant__ double c_q1_aff[6 * 36];     // affine-element tables, see q1_K_affine
__constant__ double c_q1_Nsum[8];        // sum_q N_a(q)
__constant__ double c_q1_n1[4];           // [q1d][i] 1-D linear shape values at the two Gauss abscissae
__constant__ double c_q1_N[8 * 8];       // [q][a] shape-function values at the Gauss points

struct Q1Params {
    const double* coords; const int32_t* conn; int64_t n_elems;
    const int32_t* slot;                                  // [n_elems][8][8]
    const int32_t* eqn; const uint8_t* status; const double* presc; const double* values;  // per node (ds = 1)
    double* val; double* rhs; double factor; int incremental;
};

__global__ void __launch_bounds__(128) k_q1hex_laplace(const Q1Params p) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n_elems) return;
    int32_t node[8];
    {
        const int4* c4 = reinterpret_cast<const int4*>(p.conn + e * 8);
        const int4 a = __ldg(c4), b = __ldg(c4 + 1);
        node[0] = a.x; node[1] = a.y; node[2] = a.z; node[3] = a.w;
        node[4] = b.x; node[5] = b.y; node[6] = b.z; node[7] = b.w;
    }
    double X[8][3];
#pragma unroll
    for (int a = 0; a < 8; a++) {
        const double* c = p.coords + (size_t)node[a] * 3;
        X[a][0] = __ldg(c); X[a][1] = __ldg(c + 1); X[a][2] = __ldg(c + 2);
    }
    double K[36];
#pragma unroll
    for (int k = 0; k < 36; k++) K[k] = 0.;
#pragma unroll 1
    for (int q = 0; q < 8; q++) {
        const double* dN = c_q1_dN + q * 24;
        double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int al = 0; al < 3; al++) J[i][al] = fma(X[a][i], dN[a * 3 + al], J[i][al]);
        // contra = (J^T)^{-1}: contra[r][c] = cof(J)[r][c] / det
        double co[3][3];
        co[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        co[0][1] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        co[0][2] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        co[1][0] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
        co[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
        co[1][2] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
        co[2][0] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
        co[2][1] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
        co[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        const double det = J[0][0] * co[0][0] + (J[0][1] * co[0][1] + J[0][2] * co[0][2]);
        const double id = 1.0 / det;
        const double scal = p.factor * det * c_q1_w[q];
        double g[8][3];
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int r = 0; r < 3; r++)
                g[a][r] = (co[r][0] * dN[a * 3] + co[r][1] * dN[a * 3 + 1] + co[r][2] * dN[a * 3 + 2]) * id;
        int k = 0;
#pragma unroll
        for (int a = 0; a < 8; a++) {
            const double h0 = g[a][0] * scal, h1 = g[a][1] * scal, h2 = g[a][2] * scal;
#pragma unroll
            for (int b = a; b < 8; b++, k++) K[k] = fma(h0, g[b][0], fma(h1, g[b][1], fma(h2, g[b][2], K[k])));
        }
    }
    // scatter
    const int32_t* sl = p.slot + e * 64;
    int32_t rows[8]; double gval[8]; bool cons[8]; bool any_cons = false;
#pragma unroll
    for (int a = 0; a < 8; a++) {
        rows[a] = __ldg(p.eqn + node[a]);
        cons[a] = __ldg(p.status + node[a]) == ISL_CONSTRAINED;
        any_cons |= cons[a];
        gval[a] = 0.;
    }
    if (any_cons) {
#pragma unroll
        for (int a = 0; a < 8; a++)
            if (cons[a]) gval[a] = p.incremental ? p.presc[node[a]] - p.values[node[a]] : p.presc[node[a]];
    }
    int k = 0;
#pragma unroll
    for (int a = 0; a < 8; a++) {
        const int4 s0 = __ldg(reinterpret_cast<const int4*>(sl + a * 8));
        const int4 s1 = __ldg(reinterpret_cast<const int4*>(sl + a * 8 + 4));
        const int32_t srow[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        double lift = 0.;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            // symmetric storage index of (min,max)
            const int lo = a < b ? a : b, hi = a < b ? b : a;
            const int idx = lo * 8 - (lo * (lo - 1)) / 2 + (hi - lo);
            const double v = K[idx];
            if (srow[b] >= 0) atomicAdd(p.val + srow[b], v);
            else if (cons[b]) lift += gval[b] * v;
        }
        if (any_cons && rows[a] >= 0 && lift != 0.) atomicAdd(p.rhs + rows[a], -lift);
    }
    (void)k;
}

// ---------------------------------------------------------------------------------------------
// pattern build kernels
__global__ void k_elem_eqn(const int32_t* ed, const int32_t* eqn, int64_t n_elems, int ndpe, int ds, int32_t* out) {
    const int64_t n = n_elems * ndpe * ds;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ef = t / ds; const int c = (int)(t % ds);
        out[t] = eqn[(size_t)ed[ef] * ds + c];
    }
}
// bounding box of the nodes (ordered-integer trick for atomic min / max of doubles)
__device__ __forceinline__ unsigned long long dbl_ord(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void k_bbox(const double* coords, int64_t n_nodes, int dim, unsigned long long* mn, unsigned long long* mx) {
    unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0ull, 0ull, 0ull};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += (int64_t)gridDim.x * blockDim.x)
        for (int d = 0; d < dim; d++) {
            const unsigned long long o = dbl_ord(coords[i * dim + d]);
            lo[d] = min(lo[d], o); hi[d] = max(hi[d], o);
        }
    for (int d = 0; d < dim; d++) {   // one atomic per warp
        for (int off = 16; off > 0; off >>= 1) {
            lo[d] = min(lo[d], __shfl_down_sync(0xffffffffu, lo[d], off));
            hi[d] = max(hi[d], __shfl_down_sync(0xffffffffu, hi[d], off));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(mn + d, lo[d]); atomicMax(mx + d, hi[d]); }
    }
}
__device__ __forceinline__ double ord_dbl(unsigned long long o) {
    const unsigned long long b = (o & 0x8000000000000000ull) ? (o & 0x7fffffffffffffffull) : ~o;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ unsigned int spread3(unsigned int v) {   // 10 bits -> every third bit
    v &= 0x3ff; v = (v | (v << 16)) & 0x030000ff; v = (v | (v << 8)) & 0x0300f00f; v = (v | (v << 4)) & 0x030c30c3;
    return (v | (v << 2)) & 0x09249249;
}
// locality key of an element: Z-curve index of its centroid (30 bits)
__global__ void k_elem_morton(const double* coords, const int32_t* conn, int64_t n, int npe, int dim, const unsigned long long* mn,
                              const unsigned long long* mx, int32_t* key, int32_t* idx) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        unsigned int code = 0;
        for (int d = 0; d < dim; d++) {
            double c = 0.;
            for (int a = 0; a < npe; a++) c += coords[(size_t)conn[e * npe + a] * dim + d];
            c /= npe;
            const double lo = ord_dbl(mn[d]), hi = ord_dbl(mx[d]);
            const double t = (hi > lo) ? (c - lo) / (hi - lo) : 0.;
            const unsigned int q = (unsigned int)min(1023.0, max(0.0, t * 1024.0));
            code |= spread3(q) << d;
        }
        key[e] = (int32_t)code; idx[e] = (int32_t)e;
    }
}
__global__ void k_elem_min_eqn(const int32_t* elem_eqn, int64_t n, int per, int32_t* key, int32_t* idx) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int32_t m = 0x7fffffff;
        for (int a = 0; a < per; a++) { const int32_t q = elem_eqn[e * per + a]; if (q >= 0 && q < m) m = q; }
        key[e] = m; idx[e] = (int32_t)e;
    }
}
__global__ void k_make_keys(const int32_t* er, const int32_t* ec, int64_t n_elems, int nr, int ncl, uint64_t* keys) {
    const int64_t n = n_elems * nr * ncl;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / (nr * ncl); const int ij = (int)(t % (nr * ncl));
        const int32_t r = er[e * nr + ij / ncl], c = ec[e * ncl + ij % ncl];
        keys[t] = (r < 0 || c < 0) ? ~0ull : (((uint64_t)(uint32_t)r << 32) | (uint32_t)c);
    }
}
__global__ void k_rowptr_from_keys(const uint64_t* keys, int64_t nnz, int64_t n, int64_t* rowptr) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t target = (uint64_t)r << 32;
        int64_t lo = 0, hi = nnz;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] < target) lo = mid + 1; else hi = mid; }
        rowptr[r] = lo;
    }
}
__global__ void k_cols_from_keys(const uint64_t* keys, int64_t nnz, int32_t* col) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
        col[k] = (int32_t)(keys[k] & 0xffffffffu);
}
template <typename SLOT>
__global__ void k_slotmap(const int32_t* er, const int32_t* ec, int64_t n_elems, int nr, int ncl, const int64_t* rowptr,
                          const int32_t* col, SLOT* slot) {
    const int64_t n = n_elems * nr * ncl;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / (nr * ncl); const int ij = (int)(t % (nr * ncl));
        const int32_t r = er[e * nr + ij / ncl], c = ec[e * ncl + ij % ncl];
        slot[t] = (r < 0 || c < 0) ? (SLOT)-1 : (SLOT)find_in_row(rowptr, col, r, c);
    }
}
// node-block position map of a (test, trial) field pair: base[e][M][N] = CSR position of entry (row of (test node M,
// component 0), column of (trial node N, component 0)) when the pair is REGULAR -- all components of both nodes ACTIVE and
// numbered consecutively, the rows of node M of equal length and holding the columns of node N at the same offset -- so
// that entry (M,i; N,k) sits at base + i * len[e][M] + k; -1 otherwise (the kernels then take the per-entry path).
__global__ void k_blockmap(const int32_t* ed_t, const int32_t* eqn_t, int nt, int dst, const int32_t* ed_c, const int32_t* eqn_c, int nc, int dsc,
                           int64_t n_elems, const int64_t* rowptr, const int32_t* col, int64_t* base, int32_t* len) {
    const int64_t n = n_elems * nt * nc;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / (nt * nc); const int mn = (int)(t % (nt * nc)), M = mn / nc, N = mn % nc;
        const int32_t* qm = eqn_t + (size_t)ed_t[e * nt + M] * dst;
        const int32_t* qn = eqn_c + (size_t)ed_c[e * nc + N] * dsc;
        const int32_t r0 = qm[0], c0 = qn[0];
        bool ok = r0 >= 0 && c0 >= 0;
        for (int i = 1; i < dst && ok; i++) ok = qm[i] == r0 + i;
        for (int k = 1; k < dsc && ok; k++) ok = qn[k] == c0 + k;
        int64_t b = -1;
        const int32_t L = (r0 >= 0) ? (int32_t)(rowptr[r0 + 1] - rowptr[r0]) : 0;   // (independent of the trial node: len[e][M])
        if (ok) {
            const int64_t s0 = rowptr[r0];
            for (int i = 1; i < dst && ok; i++) ok = (rowptr[r0 + i + 1] - rowptr[r0 + i]) == L;
            const int64_t p0 = ok ? find_in_row(rowptr, col, r0, c0) : -1;
            ok = ok && p0 >= 0 && p0 + dsc <= s0 + L;
            for (int i = 0; i < dst && ok; i++)
                for (int k = 0; k < dsc && ok; k++) ok = col[p0 + (int64_t)i * L + k] == c0 + k;
            if (ok) b = p0;
        }
        base[t] = b;
        if (N == 0) len[e * nt + M] = L;
    }
}
__global__ void k_remap_values(const int64_t* old_rowptr, const int32_t* old_col, const double* old_val, int64_t n,
                               const int64_t* rowptr, const int32_t* col, double* val) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        for (int64_t k = old_rowptr[r]; k < old_rowptr[r + 1]; k++) {
            const int64_t pos = find_in_row(rowptr, col, (int32_t)r, old_col[k]);
            if (pos >= 0) val[pos] = old_val[k];
        }
}
__global__ void k_insert_lhs(const double* mat, const int64_t* rows, int nr, const int64_t* cols, int ncl,
                             const int64_t* rowptr, const int32_t* col, double* val, int* err) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nr * ncl) return;
    const int64_t pos = find_in_row(rowptr, col, (int32_t)rows[t / ncl], (int32_t)cols[t % ncl]);
    if (pos < 0) { *err = 1; return; }
    atomicAdd(val + pos, mat[t]);
}
__global__ void k_insert_rhs(const double* vec, const int64_t* rows, int nr, double* rhs) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nr) atomicAdd(rhs + rows[t], vec[t]);
}
__global__ void k_count_diff(const int32_t* a, const int32_t* b, int64_t n, int* ndiff) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (a[i] != b[i]) atomicAdd(ndiff, 1);
}
__global__ void k_sumsq(const double* x, int64_t n, double* out) {
    double acc = 0.;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += x[i] * x[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
__global__ void k_pack(const double* src, const int64_t* idx, int64_t n, double* out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = src[idx[i]];
}
// base/dof/Distribute.hpp:163-210 on the device: ACTIVE components take (SET) or add (ADD) the solver's value
__global__ void k_distribute_active(const int32_t* eqn, const double* x, double* values, int64_t n, int add) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t q = eqn[k];
        if (q >= 0) values[k] = add ? values[k] + x[q] : x[q];
    }
}
__global__ void k_eqn2dof(const int32_t* eqn, int64_t n, int32_t* map) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        if (eqn[k] >= 0) map[eqn[k]] = (int32_t)k;
}
// ... then CONSTRAINED components are set to their constraint value: prescribed + sum_j weight_j * (value of master j)
__global__ void k_distribute_constrained(const uint8_t* status, const double* presc, const int32_t* cptr, const int32_t* cm,
                                         const double* cw, const int32_t* eqn2dof, int32_t eqn_lo, int32_t eqn_hi, double* values,
                                         int64_t n, int* err) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        if (status[k] != ISL_CONSTRAINED) continue;
        double v = presc[k];
        if (cptr)
            for (int32_t j = cptr[k]; j < cptr[k + 1]; j++) {
                const int32_t q = cm[j];
                const int32_t d = (q >= eqn_lo && q < eqn_hi) ? eqn2dof[q - eqn_lo] : -1;
                if (d < 0) { *err = 1; continue; }   // master in another field: not supported on the device
                v = fma(cw[j], values[d], v);
            }
        values[k] = v;
    }
}
__global__ void k_unpack_add(double* dst, const int64_t* idx, int64_t n, const double* in) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(dst + idx[i], in[i]);
}

#include "isl_comm.cuh"
#include "isl_patch.cuh"
#include "isl_rowgather.cuh"
#include "isl_rows_fromk.cuh"
#include "isl_patch_dev.cuh"

// ---------------------------------------------------------------------------------------------
struct FieldDev {
    bool set = false;
    bool dof_is_node = false;  // elem_dof identical to the mesh connectivity (isoparametric copy)
    int deg = 0, ds = 0, ndpe = 0; int64_t n_obj = 0;
    DevBuf<int32_t> elem_dof, eqn; DevBuf<uint8_t> status; DevBuf<double> presc, values;
    DevBuf<int32_t> elem_eqn;  // [n_elems][ndpe*ds]
    DevBuf<int32_t> eorder;    // locality-sorted element order of the generic kernels (get_eorder)
    // linear constraints with master DoFs (isl_field_set_constraints): dense pointer array [n_obj*ds+1], masters as
    // equation numbers; host copies feed the extra pattern keys in build_pattern
    bool has_masters = false;
    DevBuf<int32_t> cptr, cmaster; DevBuf<double> cweight;
    std::vector<int32_t> h_cptr, h_cmaster, h_elem_dof, h_eqn;
    void reset_constraints() {
        has_masters = false; cptr.release(); cmaster.release(); cweight.release();
        h_cptr.clear(); h_cmaster.clear();
    }
    void reset() {
        set = false; dof_is_node = false; deg = ds = ndpe = 0; n_obj = 0;
        elem_dof.release(); eqn.release(); status.release(); presc.release(); values.release(); elem_eqn.release(); eorder.release();
        reset_constraints(); h_elem_dof.clear(); h_eqn.clear();
    }
};

struct TableDev {
    int nq = 0;
    DevBuf<double> w, dNg, Ng, Nt, dNt, Nc, dNc;
};

}  // namespace

struct isl_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    int n_sm = 148;
    int64_t launches = 0;
    // mesh
    int shape = 0, geom_deg = 0, dim = 0, npe = 0; int64_t n_nodes = 0, n_elems = 0;
    int64_t n_owned = 0;  // elements [0, n_owned) are assembled here; the rest (halo) only shape the pattern
    DevBuf<double> coords; DevBuf<int32_t> conn;
    FieldDev fields[5];
    // system
    int64_t n_eqn = -1, nnz = 0;
    DevBuf<int64_t> rowptr; DevBuf<int32_t> col; DevBuf<double> val, rhs;
    // isl_get_csr_async: second set of value / rhs buffers and a copy stream, so that the device -> host copy of a finished
    // system overlaps the assembly (and the host -> device copies) of the next one
    DevBuf<double> val_alt, rhs_alt; cudaStream_t copy_stream = nullptr; cudaEvent_t ev_asm = nullptr, ev_copy = nullptr;
    bool copy_in_flight = false, sys_released = false;
    std::set<std::pair<int, int>> pattern_pairs, sys_pairs;
    struct SlotMap { DevBuf<int32_t> s32; DevBuf<int64_t> s64; };
    std::map<std::pair<int, int>, std::unique_ptr<SlotMap>> slotmaps;
    std::map<int, std::unique_ptr<GatherSet>> gathersets;   // per field: tables of the atomic-free hyperelastic path
    int hypel_gather = 1;      // ISL_HYPEL_GATHER
    // atomic-free path of the generic kernels (isl_gather.cuh, second half): (element, row) pairs per test field, column
    // positions per (test, trial, compact), one scratch buffer for the element matrices of the operation in flight
    std::map<int, std::unique_ptr<RowPairs>> rowpairs;
    std::map<std::array<int, 3>, std::unique_ptr<GenGatherSet>> gengathers;
    DevBuf<double> gen_kbuf;
    int gen_gather = 1;        // ISL_GEN_GATHER (0: atomic scatter through the slot / node-block maps)
    int gen_range = 1;         // ISL_GEN_RANGE: row buffers cover the column range of the block only
    int gen_sampled = 1;       // ISL_GEN_SAMPLED: isl_assemble_matrix_sampled through the atomic-free path too (0: atomic scatter)
    int gen_krow = 1;          // ISL_GEN_KROW: the gather reads the element-matrix row of a pair from a table (0: two integer divisions per pair)
    int gen_tile = 1;          // ISL_GEN_TILE: k_tangent_strips for the Laplace-type integrands and the Stokes coupling blocks (0: per-entry loops of k_tangent)
    int hypel_mc_small = 5;    // ISL_HYPEL_MC: tile height for elements with at most 10 nodes (2, 3 or 5)
    int hypel_occ3 = 1;        // ISL_HYPEL_OCC3: Q2 variant of the tile kernel compiled for three CTAs per SM
    struct BlockMap { DevBuf<int64_t> base; DevBuf<int32_t> len; };
    std::map<std::pair<int, int>, std::unique_ptr<BlockMap>> blockmaps;   // per (test, trial) pair
    int block_slots = 1;       // ISL_BLOCK_SLOTS: 0 per-entry slot maps only, 1 node-block maps in the generic kernels, 2 also in the tile kernel
    bool wide_slots = false;   // nnz >= 2^31 (or ISL_SLOT64=1): CSR positions do not fit 32 bits
    bool force_slot64 = false;
    int fromk_tile_order = 0;  // row tiles of the general Q1 path along a Z-curve (ISL_FROMK_TILE_ORDER=1): measured slower, 4.00 vs 3.67 ms (session27)
    int stage_kb = 48;         // staging budget of the generic kernels per CTA (ISL_STAGE_KB; sweep in profiles/r2/session23.log)
    std::map<std::array<int, 3>, std::unique_ptr<TableDev>> tables;  // (quad_deg, test, trial)
    bool q1_tables_loaded = false;
    DevBuf<double> scratch_d; DevBuf<int> scratch_i;
    // patch assembly (isl_patch.cuh)
    std::map<int, std::unique_ptr<PatchSet>> patchsets;  // per field
    std::map<int, std::unique_ptr<FromKSet>> fromk_sets; // per field: two-kernel path for general (non-affine) Q1 elements
    bool val_is_zero = false;   // matrix values known to be zero (fresh solver): complete rows may be stored
    bool val_zero_pending = false;  // the memset of val has been postponed (a store-mode patch launch makes it unnecessary)
    struct PendingQ1 { bool active = false; int field = 0; double factor = 1.; int incremental = 1; } pending_q1;
    int q1_mode = 1;            // 0 = one thread per element + atomics, 1 = shared-memory patches
    int patch_rows = 256, patch_threads = 128, patch_ctas_per_sm = 2;
    double patch_stretch = 3.0;  // ISL_PATCH_STRETCH: boxes this many times longer along the axis of consecutive equation numbers
    int q1_fast = 3;            // bit0: sum-factorised local matrix, bit1: affine-element shortcut
    int affine_kernel = 1;      // all-affine meshes: low-register kernel with 256 threads per CTA
    int q1_rows = 1;            // row kernels (isl_rowgather.cuh: all-affine meshes; isl_rows_fromk.cuh: general elements);
                                // ISL_Q1_ROWS=0 selects the round-1 shared-memory patch kernels
    int rows_threads = 128;     // CTA size of the affine row kernel (ISL_ROWS_THREADS)
    int fromk_pipeline = 0;     // ISL_FROMK_PIPELINE=1: general Q1 elements in one persistent producer/consumer kernel (measured slower: 6.5 vs 3.8 ms)
    int64_t pipe_rows = 65536;  // rows per pipeline chunk (ISL_PIPE_ROWS)
    int fromk_variant = 0;      // row-kernel variant of the two-kernel general path (ISL_FROMK_VARIANT)
    int aff_split = 1;          // mbarrier arrive/wait phases in the all-affine kernel (ISL_AFF_SPLIT=0: __syncthreads)
    int patch_threads_aff = 0;  // experiment knob: alternative CTA size of the all-affine kernel
    int affine_state = -1;      // -1 unknown, 0 some element is not affine, 1 every owned element is affine
    int patch_ws = 0;           // warp-specialised patch kernel (compute warps + scatter warps, one CTA per SM)
    int defer_launch = 1;       // fuse stiffness + body force of the Q1 hot path into one launch
    DevBuf<unsigned long long> profbuf;
    std::unique_ptr<CommState> comm;   // NCCL communicator + interface-row exchange plan (isl_comm.cuh)
    DevBuf<int64_t> l2g;               // local -> global equation (exchange plan)
    bool sys_stale = false;     // mesh or field arrays were replaced while the system held matrix entries: they are gone
    int tangent_sym = 1;        // symmetric register-tiled hyperelastic tangent (k_tangent_hypel_sym); ISL_TANGENT_SYM=0: k_tangent
    int elem_order = 1;         // generic kernels walk the elements in a locality-sorted order (ISL_ELEM_ORDER=0: as given)
    int tangent_tiled = 0;      // ISL_TANGENT_TILED=1: register-tiled hyperelastic tangent (isl_tangent_tiled.cuh; not yet default)

    // staging of isl_insert_lhs / isl_insert_rhs operands: a ring of slices so that a copy never overwrites operands a
    // queued kernel still reads only after the stream has passed them (stream order) -- one arena, grown on demand
    DevBuf<char> ins_buf; size_t ins_off = 0; DevBuf<int> ins_err;
    char* insert_arena(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        if (!ins_err.p) { ins_err.alloc(1); cudaMemsetAsync(ins_err.p, 0, sizeof(int), stream); }
        if (bytes > ins_buf.n || !ins_buf.p) {
            cudaStreamSynchronize(stream);
            ins_buf.alloc(std::max<size_t>(bytes * 4, (size_t)1 << 20)); ins_off = 0;
        }
        if (ins_off + bytes > ins_buf.n) ins_off = 0;   // wrap: everything in the arena is consumed in stream order
        char* p = ins_buf.p + ins_off; ins_off += bytes;
        return p;
    }

    int grid_for(int64_t n, int block) const {
        const int64_t g = (n + block - 1) / block;
        return (int)std::max<int64_t>(1, std::min<int64_t>(g, (int64_t)n_sm * 32));
    }
};

namespace {

#define ISL_LAUNCH(eng, kernel, grid, block, smem, ...)                       \
    do {                                                                      \
        kernel<<<(grid), (block), (smem), (eng)->stream>>>(__VA_ARGS__);      \
        (eng)->launches++;                                                    \
        ISL_CUDA(cudaGetLastError());                                         \
    } while (0)

void materialize_zero(isl_engine* h);
void comm_join(isl_engine* h);
void flush_pending(isl_engine* h, int fuse_body, double f0);
inline void flush_pending(isl_engine* h) { flush_pending(h, 0, 0.); }

template <class T>
void upload(isl_engine* h, DevBuf<T>& dst, const T* src, size_t n) {
    dst.alloc(n);
    if (n) ISL_CUDA(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyDefault, h->stream));
}
template <class T>
void upload_vec(isl_engine* h, DevBuf<T>& dst, const std::vector<T>& v) {
    dst.alloc(v.size());
    if (!v.empty()) {
        ISL_CUDA(cudaMemcpyAsync(dst.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));  // v may be a temporary
    }
}

// replacing mesh / field arrays drops the pattern and with it the matrix values: if the current system already holds
// matrix contributions it becomes unusable (a second FieldBinder on the same solver, a binder rebuilt between two assembly
// calls); later calls on it fail instead of silently continuing on an empty matrix
void mark_stale_if_touched(isl_engine* h) { if (!h->val_is_zero && h->nnz > 0) h->sys_stale = true; }   // rhs survives, the matrix does not
void require_live_system(isl_engine* h) {
    ISL_REQUIRE(!h->sys_stale, "mesh or field arrays were replaced after assembly into this system had started, so its entries "
                               "are gone: create a new solver first (one FieldBinder per solver is supported)");
    ISL_REQUIRE(!h->sys_released, "the system was handed to the host by isl_get_csr_async: create a new solver first");
    if (h->comm) h->comm->iface_event_valid = false;   // set again by the split launch of the Q1 row kernel only
}

void invalidate_pattern(isl_engine* h) {
    h->pattern_pairs.clear();
    h->slotmaps.clear(); h->blockmaps.clear(); h->gathersets.clear(); h->rowpairs.clear(); h->gengathers.clear();
    h->patchsets.clear();
    h->fromk_sets.clear();
    h->nnz = 0;
    h->rowptr.release(); h->col.release(); h->val.release();
    if (h->copy_in_flight && h->ev_copy) cudaEventSynchronize(h->ev_copy);
    h->val_alt.release(); h->copy_in_flight = false;
}

void build_elem_eqn(isl_engine* h, FieldDev& f) {
    if (f.elem_eqn.p) return;
    const int64_t n = h->n_elems * f.ndpe * f.ds;
    f.elem_eqn.alloc(n);
    ISL_LAUNCH(h, k_elem_eqn, h->grid_for(n, 256), 256, 0, f.elem_dof.p, f.eqn.p, h->n_elems, f.ndpe, f.ds, f.elem_eqn.p);
}

// rebuild CSR for `pairs`; existing values are carried over into the new layout
void build_pattern(isl_engine* h, const std::set<std::pair<int, int>>& pairs) {
    ISL_REQUIRE(h->n_eqn >= 0, "isl_system_create must be called first");
    ISL_REQUIRE(h->n_eqn < (int64_t)1 << 31, "more than 2^31 equations are not supported");
    // values of the old layout are carried over -- unless there are none yet (pattern registration before the first
    // assembly call): then the old arrays are released first, their memory is needed for the keys of large systems
    const bool carry = h->nnz > 0 && h->val.p && !h->val_is_zero;
    if (carry) materialize_zero(h);
    else { h->val.release(); h->col.release(); h->slotmaps.clear(); h->blockmaps.clear(); h->gathersets.clear(); h->rowpairs.clear(); h->gengathers.clear(); h->nnz = 0; h->val_zero_pending = false; }
    int64_t total = 0;
    for (auto& pr : pairs) {
        FieldDev& t = h->fields[pr.first]; FieldDev& c = h->fields[pr.second];
        ISL_REQUIRE(t.set && c.set, "pattern registration on a field that was not set");
        total += h->n_elems * (int64_t)(t.ndpe * t.ds) * (c.ndpe * c.ds);
    }
    // pairs whose test or trial field has slaves of master DoFs reach further entries: effective rows x effective
    // columns (ACTIVE ids ++ master ids) of every element holding such a slave (solver/TripletContainer.hpp:229-262);
    // few, formed on the host
    std::vector<uint64_t> extra;
    for (auto& pr : pairs) {
        FieldDev& t = h->fields[pr.first]; FieldDev& c = h->fields[pr.second];
        if (!t.has_masters && !c.has_masters) continue;
        auto host_copy = [&](FieldDev& f) {
            if (!f.h_elem_dof.empty()) return;
            f.h_elem_dof.resize((size_t)h->n_elems * f.ndpe); f.h_eqn.resize((size_t)f.n_obj * f.ds);
            ISL_CUDA(cudaMemcpyAsync(f.h_elem_dof.data(), f.elem_dof.p, f.h_elem_dof.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            ISL_CUDA(cudaMemcpyAsync(f.h_eqn.data(), f.eqn.p, f.h_eqn.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            ISL_CUDA(cudaStreamSynchronize(h->stream));
        };
        host_copy(t); host_copy(c);
        auto effective = [&](const FieldDev& f, int64_t e, std::vector<int32_t>& eff) {
            return isl_effective_ids(f.h_elem_dof.data(), f.ndpe, f.ds, f.h_eqn.data(), f.has_masters ? f.h_cptr.data() : nullptr,
                                     f.h_cmaster.data(), e, eff);
        };
        std::vector<int32_t> er, ec;
        for (int64_t e = 0; e < h->n_elems; e++) {
            const bool sr = effective(t, e, er), sc = effective(c, e, ec);
            if (!sr && !sc) continue;
            for (int32_t r : er) for (int32_t cc : ec) extra.push_back(((uint64_t)(uint32_t)r << 32) | (uint32_t)cc);
        }
    }
    total += (int64_t)extra.size();
    DevBuf<uint64_t> keys, keys2;
    keys.alloc(total + 1);
    int64_t off = 0;
    if (!extra.empty()) {
        ISL_CUDA(cudaMemcpyAsync(keys.p, extra.data(), extra.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        off = (int64_t)extra.size();
    }
    for (auto& pr : pairs) {
        FieldDev& t = h->fields[pr.first]; FieldDev& c = h->fields[pr.second];
        build_elem_eqn(h, t); build_elem_eqn(h, c);
        const int nr = t.ndpe * t.ds, ncl = c.ndpe * c.ds;
        const int64_t n = h->n_elems * nr * ncl;
        ISL_LAUNCH(h, k_make_keys, h->grid_for(n, 256), 256, 0, t.elem_eqn.p, c.elem_eqn.p, h->n_elems, nr, ncl, keys.p + off);
        off += n;
    }
    // sentinel so that the invalid key always exists exactly as the last unique key
    const uint64_t inval = ~0ull;
    ISL_CUDA(cudaMemcpyAsync(keys.p + total, &inval, sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    total += 1;
    keys2.alloc(total);
    // the invalid key has all bits set: sort on the full 64 bits so it stays last
    size_t tmp_bytes = 0;
    ISL_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys.p, keys2.p, total, 0, 64, h->stream));
    DevBuf<char> tmp; tmp.alloc(tmp_bytes);
    ISL_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, keys.p, keys2.p, total, 0, 64, h->stream));
    h->launches += 8;
    DevBuf<int64_t> nsel; nsel.alloc(1);
    size_t tmp2 = 0;
    ISL_CUDA(cub::DeviceSelect::Unique(nullptr, tmp2, keys2.p, keys.p, nsel.p, total, h->stream));
    if (tmp2 > tmp.n) tmp.alloc(tmp2);
    ISL_CUDA(cub::DeviceSelect::Unique(tmp.p, tmp2, keys2.p, keys.p, nsel.p, total, h->stream));
    h->launches += 2;
    int64_t nuniq = 0;
    ISL_CUDA(cudaMemcpyAsync(&nuniq, nsel.p, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    ISL_CUDA(cudaStreamSynchronize(h->stream));
    const int64_t nnz = nuniq - 1;  // drop the invalid key
    h->wide_slots = nnz >= ((int64_t)1 << 31) || h->force_slot64;
    keys2.release(); tmp.release();

    DevBuf<int64_t> rowptr; DevBuf<int32_t> col; DevBuf<double> val;
    rowptr.alloc(h->n_eqn + 1); col.alloc(nnz); val.alloc(nnz);
    ISL_LAUNCH(h, k_rowptr_from_keys, h->grid_for(h->n_eqn + 1, 256), 256, 0, keys.p, nnz, h->n_eqn, rowptr.p);
    if (nnz) {
        ISL_LAUNCH(h, k_cols_from_keys, h->grid_for(nnz, 256), 256, 0, keys.p, nnz, col.p);
        ISL_CUDA(cudaMemsetAsync(val.p, 0, nnz * sizeof(double), h->stream));
    }
    if (carry)
        ISL_LAUNCH(h, k_remap_values, h->grid_for(h->n_eqn, 128), 128, 0, h->rowptr.p, h->col.p, h->val.p, h->n_eqn,
                   rowptr.p, col.p, val.p);
    ISL_CUDA(cudaStreamSynchronize(h->stream));
    h->rowptr.swap(rowptr); h->col.swap(col); h->val.swap(val);
    h->nnz = nnz;
    h->pattern_pairs = pairs;
    h->slotmaps.clear(); h->blockmaps.clear(); h->gathersets.clear(); h->rowpairs.clear(); h->gengathers.clear();
    h->patchsets.clear();
    h->fromk_sets.clear();
}

void ensure_pair(isl_engine* h, int t, int c) {
    ISL_REQUIRE(t >= 0 && t < 5 && c >= 0 && c < 5, "field index out of range");
    h->sys_pairs.insert({t, c});
    if (!h->pattern_pairs.count({t, c})) {
        auto pairs = h->pattern_pairs;
        pairs.insert({t, c});
        build_pattern(h, pairs);
    }
}

const isl_engine::SlotMap* get_slotmap(isl_engine* h, int t, int c) {
    auto key = std::make_pair(t, c);
    auto it = h->slotmaps.find(key);
    if (it != h->slotmaps.end()) return it->second.get();
    FieldDev& ft = h->fields[t]; FieldDev& fc = h->fields[c];
    build_elem_eqn(h, ft); build_elem_eqn(h, fc);
    const int nr = ft.ndpe * ft.ds, ncl = fc.ndpe * fc.ds;
    const int64_t n = h->n_owned * nr * ncl;
    auto buf = std::make_unique<isl_engine::SlotMap>();
    if (h->wide_slots) {
        buf->s64.alloc(n);
        ISL_LAUNCH(h, k_slotmap<int64_t>, h->grid_for(n, 256), 256, 0, ft.elem_eqn.p, fc.elem_eqn.p, h->n_owned, nr, ncl,
                   h->rowptr.p, h->col.p, buf->s64.p);
    } else {
        buf->s32.alloc(n);
        ISL_LAUNCH(h, k_slotmap<int32_t>, h->grid_for(n, 256), 256, 0, ft.elem_eqn.p, fc.elem_eqn.p, h->n_owned, nr, ncl,
                   h->rowptr.p, h->col.p, buf->s32.p);
    }
    const isl_engine::SlotMap* p = buf.get();
    h->slotmaps[key] = std::move(buf);
    return p;
}
void bind_slots(isl_engine* h, AsmParams& p, int t, int c) {
    const isl_engine::SlotMap* m = get_slotmap(h, t, c);
    p.slot = m->s32.p; p.slot64 = m->s64.p;
}
bool bind_blockmap(isl_engine* h, AsmParams& p, int t, int c) {
    FieldDev& ft = h->fields[t]; FieldDev& fc = h->fields[c];
    if (!h->block_slots || ft.ds * fc.ds <= 1 || ft.ds > 3 || fc.ds > 3 || ft.has_masters || fc.has_masters) return false;
    auto key = std::make_pair(t, c);
    auto it = h->blockmaps.find(key);
    if (it == h->blockmaps.end()) {
        auto bm = std::make_unique<isl_engine::BlockMap>();
        const int64_t n = h->n_owned * ft.ndpe * fc.ndpe;
        bm->base.alloc(n); bm->len.alloc(h->n_owned * ft.ndpe);
        ISL_LAUNCH(h, k_blockmap, h->grid_for(n, 256), 256, 0, ft.elem_dof.p, ft.eqn.p, ft.ndpe, ft.ds, fc.elem_dof.p, fc.eqn.p, fc.ndpe, fc.ds,
                   h->n_owned, h->rowptr.p, h->col.p, bm->base.p, bm->len.p);
        it = h->blockmaps.emplace(key, std::move(bm)).first;
    }
    p.bbase = it->second->base.p; p.blen = it->second->len.p;
    return true;
}

TableDev* get_tables(isl_engine* h, int quad_deg, int t, int c) {
    std::array<int, 3> key = {quad_deg, t, c};
    auto it = h->tables.find(key);
    if (it != h->tables.end()) return it->second.get();
    const isl::Rule R = isl::make_rule(h->shape, quad_deg);
    const isl::Basis G(h->shape, h->geom_deg), T(h->shape, h->fields[t].deg), C(h->shape, h->fields[c].deg);
    auto td = std::make_unique<TableDev>();
    td->nq = R.n;
    auto tab = [&](const isl::Basis& B, std::vector<double>& N, std::vector<double>& dN) {
        N.resize((size_t)R.n * B.nfun); dN.resize((size_t)R.n * B.nfun * B.dim);
        for (int q = 0; q < R.n; q++) B.eval(&R.p[(size_t)q * R.dim], &N[(size_t)q * B.nfun], &dN[(size_t)q * B.nfun * B.dim]);
    };
    std::vector<double> N, dN;
    upload_vec(h, td->w, R.w);
    tab(G, N, dN); upload_vec(h, td->Ng, N); upload_vec(h, td->dNg, dN);
    tab(T, N, dN); upload_vec(h, td->Nt, N); upload_vec(h, td->dNt, dN);
    tab(C, N, dN); upload_vec(h, td->Nc, N); upload_vec(h, td->dNc, dN);
    TableDev* p = td.get();
    h->tables[key] = std::move(td);
    return p;
}

size_t stage_doubles_per_elem(const AsmParams& p, int dim) {
    size_t n = (size_t)p.npe * dim + (size_t)p.nq * dim * dim + p.nq;
    if (p.need_gt) n += (size_t)p.nq * p.nt * dim;
    if (p.need_gc && !(p.bubnov && p.need_gt)) n += (size_t)p.nq * p.nc * dim;
    n += (size_t)p.nq * p.nqdata;
    return n;
}

const int32_t* get_eorder(isl_engine* h, int field);

void fill_common(isl_engine* h, AsmParams& p, int quad_deg, int t, int c) {
    ISL_REQUIRE(h->n_elems > 0, "mesh not set");
    FieldDev& ft = h->fields[t]; FieldDev& fc = h->fields[c];
    ISL_REQUIRE(ft.set && fc.set, "field not set");
    TableDev* td = get_tables(h, quad_deg, t, c);
    p.coords = h->coords.p; p.conn = h->conn.p; p.n_elems = h->n_owned; p.npe = h->npe;
    p.w = td->w.p; p.dNg = td->dNg.p; p.Ng = td->Ng.p; p.Nt = td->Nt.p; p.dNt = td->dNt.p; p.Nc = td->Nc.p; p.dNc = td->dNc.p;
    p.nq = td->nq; p.nt = ft.ndpe; p.nc = fc.ndpe; p.dst = ft.ds; p.dsc = fc.ds; p.bubnov = (t == c);
    p.ed_t = ft.elem_dof.p; p.ed_c = fc.elem_dof.p; p.eqn_t = ft.eqn.p; p.eqn_c = fc.eqn.p;
    p.st_c = fc.status.p; p.presc_c = fc.presc.p; p.val_c = fc.values.p;
    p.val = h->val.p; p.rhs = h->rhs.p;
    p.cptr_t = ft.has_masters ? ft.cptr.p : nullptr; p.cm_t = ft.cmaster.p; p.cw_t = ft.cweight.p;
    p.cptr_c = fc.has_masters ? fc.cptr.p : nullptr; p.cm_c = fc.cmaster.p; p.cw_c = fc.cweight.p;
    p.rowptr = h->rowptr.p; p.col = h->col.p;
    p.eorder = get_eorder(h, t);
}

// locality-sorted element order of a field (smallest equation number of the element), built once per numbering
const int32_t* get_eorder(isl_engine* h, int field) {
    if (!h->elem_order) return nullptr;
    FieldDev& f = h->fields[field];
    if (f.eorder.p && (int64_t)f.eorder.n == h->n_owned) return f.eorder.p;
    const int64_t n = h->n_owned;
    if (n <= 0 || n >= ((int64_t)1 << 31)) return nullptr;
    DevBuf<int32_t> key, key2, idx;
    key.alloc(n); key2.alloc(n); idx.alloc(n); f.eorder.alloc(n);
    // Z-curve of the element centroids: spatial neighbours are processed close in time whatever the caller's element
    // order and DoF numbering are, so the CSR rows they share are still in L2 when the next element adds to them
    DevBuf<unsigned long long> bb; bb.alloc(6);
    ISL_CUDA(cudaMemsetAsync(bb.p, 0xff, 3 * sizeof(unsigned long long), h->stream));
    ISL_CUDA(cudaMemsetAsync(bb.p + 3, 0, 3 * sizeof(unsigned long long), h->stream));
    ISL_LAUNCH(h, k_bbox, std::min(h->grid_for(h->n_nodes, 256), h->n_sm * 8), 256, 0, h->coords.p, h->n_nodes, h->dim, bb.p, bb.p + 3);
    ISL_LAUNCH(h, k_elem_morton, h->grid_for(n, 256), 256, 0, h->coords.p, h->conn.p, n, h->npe, h->dim, bb.p, bb.p + 3, key.p, idx.p);
    size_t tb = 0;
    ISL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key2.p, idx.p, f.eorder.p, n, 0, 32, h->stream));
    DevBuf<char> tmp; tmp.alloc(tb);
    ISL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, key2.p, idx.p, f.eorder.p, n, 0, 32, h->stream));
    h->launches += 4;
    ISL_CUDA(cudaStreamSynchronize(h->stream));
    return f.eorder.p;
}

GatherSet* get_gatherset(isl_engine* h, int t) {
    auto it = h->gathersets.find(t);
    if (it != h->gathersets.end()) return it->second->ok ? it->second.get() : nullptr;
    auto gs = std::make_unique<GatherSet>();
    FieldDev& f = h->fields[t];
    build_elem_eqn(h, f);
    const int nr = f.ndpe * f.ds;
    const int64_t P = h->n_owned * nr, n_rows = h->n_eqn;
    gs->nr = nr; gs->n_rows = n_rows;
    if (P > 0 && P < ((int64_t)1 << 31) && n_rows > 0) {
        DevBuf<int32_t> key, key2, idx;
        key.alloc(P); key2.alloc(P); idx.alloc(P); gs->pair.alloc(P);
        ISL_LAUNCH(h, k_gs_keys, h->grid_for(P, 256), 256, 0, f.elem_eqn.p, P, key.p, idx.p);
        size_t tb = 0;
        ISL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key2.p, idx.p, gs->pair.p, P, 0, 31, h->stream));
        DevBuf<char> tmp; tmp.alloc(tb);
        ISL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, key2.p, idx.p, gs->pair.p, P, 0, 31, h->stream));
        h->launches += 4;
        gs->row_start.alloc(n_rows + 1);
        ISL_LAUNCH(h, k_gs_row_start, h->grid_for(n_rows + 1, 256), 256, 0, key2.p, P, n_rows, gs->row_start.p);
        DevBuf<int> st; st.alloc(2);
        ISL_CUDA(cudaMemsetAsync(st.p, 0, 2 * sizeof(int), h->stream));
        ISL_LAUNCH(h, k_gs_max_len, std::min(h->grid_for(n_rows, 256), h->n_sm * 8), 256, 0, h->rowptr.p, n_rows, st.p + 1);
        ISL_LAUNCH(h, k_gs_dup, h->grid_for(P, 256), 256, 0, f.elem_eqn.p, h->n_owned, nr, st.p);
        int64_t n_pairs = 0;
        ISL_CUDA(cudaMemcpyAsync(&n_pairs, gs->row_start.p + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        gs->n_pairs = n_pairs;
        gs->pos.alloc((size_t)std::max<int64_t>(n_pairs, 1) * nr);
        if (n_pairs > 0)
            ISL_LAUNCH(h, k_gs_pos, h->grid_for(n_pairs * nr, 256), 256, 0, gs->pair.p, key2.p, n_pairs, nr, f.elem_eqn.p, h->rowptr.p, h->col.p,
                       gs->pos.p, st.p);
        int hst[2] = {0, 0};
        ISL_CUDA(cudaMemcpyAsync(hst, st.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        gs->max_len = hst[1];
        // a row buffer per warp must fit: 8 warps x max_len doubles within the shared memory of a CTA
        gs->ok = hst[0] == 0 && (size_t)gs->max_len * 8 * sizeof(double) <= 200 * 1024;
        if (gs->ok) gs->Kbuf.alloc((size_t)h->n_owned * nr * nr);
        if (getenv("ISL_VERBOSE"))
            fprintf(stderr, "[isl] atomic-free hyperelastic path: %s, %lld (element, row) pairs, longest row %d, K buffer %.2f GB, positions %.2f GB\n",
                    gs->ok ? "ok" : "not eligible", (long long)n_pairs, gs->max_len, gs->ok ? (double)h->n_owned * nr * nr * 8 / 1e9 : 0.,
                    (double)n_pairs * nr * 2 / 1e9);
    }
    GatherSet* out = gs->ok ? gs.get() : nullptr;
    h->gathersets[t] = std::move(gs);
    return out;
}

// rows gathered from the stored element matrices (isl_gather.cuh, kernel B)
void launch_gather_rows(isl_engine* h, GatherSet* gs, const AsmParams& a, const FieldDev& f, bool store) {
    GatherParams g; std::memset(&g, 0, sizeof(g));
    g.pair = gs->pair.p; g.row_start = gs->row_start.p; g.pos = gs->pos.p; g.Kbuf = gs->Kbuf.p;
    g.nr = gs->nr; g.nt = f.ndpe; g.ds = f.ds; g.n_rows = gs->n_rows;
    g.rowptr = h->rowptr.p; g.val = h->val.p; g.rhs = h->rhs.p;
    g.ed = f.elem_dof.p; g.status = f.status.p; g.presc = f.presc.p; g.values = f.values.p; g.incremental = a.incremental;
    g.store = store ? 1 : 0;
    g.buf_len = (gs->max_len + 1) & ~1;
    constexpr int W = 8;
    const size_t smem = (size_t)W * g.buf_len * sizeof(double);
    ISL_CUDA(cudaFuncSetAttribute(k_gather_rows<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    const int64_t nb = (gs->n_rows + W - 1) / W;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nb, (int64_t)h->n_sm * 16));
    k_gather_rows<W><<<grid, W * 32, smem, h->stream>>>(g);
    h->launches++;
    ISL_CUDA(cudaGetLastError());
}

// ---- atomic-free path of the generic kernels (isl_gather.cuh, second half) ----
RowPairs* get_rowpairs(isl_engine* h, int t) {
    auto it = h->rowpairs.find(t);
    if (it != h->rowpairs.end()) return it->second->ok ? it->second.get() : nullptr;
    auto rp = std::make_unique<RowPairs>();
    FieldDev& f = h->fields[t];
    build_elem_eqn(h, f);
    const int nr = f.ndpe * f.ds;
    const int64_t P = h->n_owned * nr, n_rows = h->n_eqn;
    rp->nr = nr; rp->n_rows = n_rows;
    if (P > 0 && P < ((int64_t)1 << 31) && n_rows > 0) {
        DevBuf<int32_t> key, key2, idx;
        key.alloc(P); key2.alloc(P); idx.alloc(P); rp->pair.alloc(P);
        ISL_LAUNCH(h, k_gs_keys, h->grid_for(P, 256), 256, 0, f.elem_eqn.p, P, key.p, idx.p);
        size_t tb = 0;
        ISL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key2.p, idx.p, rp->pair.p, P, 0, 31, h->stream));
        DevBuf<char> tmp; tmp.alloc(tb);
        ISL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, key2.p, idx.p, rp->pair.p, P, 0, 31, h->stream));
        h->launches += 4;
        rp->row_start.alloc(n_rows + 1);
        ISL_LAUNCH(h, k_gs_row_start, h->grid_for(n_rows + 1, 256), 256, 0, key2.p, P, n_rows, rp->row_start.p);
        int64_t n_pairs = 0;
        ISL_CUDA(cudaMemcpyAsync(&n_pairs, rp->row_start.p + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        rp->n_pairs = n_pairs;
        rp->ok = true;
    }
    RowPairs* out = rp->ok ? rp.get() : nullptr;
    h->rowpairs[t] = std::move(rp);
    return out;
}

// integrands whose local matrix repeats one scalar per node pair on every DoF component
bool gen_gather_compact(int kid) { return kid == ISL_K_LAPLACE || kid == ISL_K_VECTOR_LAPLACE || kid == ISL_K_MASS || kid == ISL_K_CONVECTION; }

GenGatherSet* get_gengather(isl_engine* h, int t, int c, int compact) {
    FieldDev& ft = h->fields[t]; FieldDev& fc = h->fields[c];
    if (ft.has_masters || fc.has_masters) return nullptr;
    if (compact && ft.ds != fc.ds) return nullptr;
    const std::array<int, 3> key = {t, c, compact};
    auto it = h->gengathers.find(key);
    if (it != h->gengathers.end()) return it->second->ok ? it->second.get() : nullptr;
    auto gs = std::make_unique<GenGatherSet>();
    gs->compact = compact;
    gs->KR = compact ? ft.ndpe : ft.ndpe * ft.ds;
    gs->KC = compact ? fc.ndpe : fc.ndpe * fc.ds;
    RowPairs* rp = gs->KC <= 96 ? get_rowpairs(h, t) : nullptr;
    if (rp) {
        build_elem_eqn(h, fc);
        const int ncl = fc.ndpe * fc.ds;
        DevBuf<int> st; st.alloc(2);
        ISL_CUDA(cudaMemsetAsync(st.p, 0, 2 * sizeof(int), h->stream));
        ISL_LAUNCH(h, k_gs_max_len, std::min(h->grid_for(rp->n_rows, 256), h->n_sm * 8), 256, 0, h->rowptr.p, rp->n_rows, st.p + 1);
        ISL_LAUNCH(h, k_gs_dup, h->grid_for(h->n_owned * ncl, 256), 256, 0, fc.elem_eqn.p, h->n_owned, ncl, st.p);
        gs->pos.alloc((size_t)std::max<int64_t>(rp->n_pairs, 1) * gs->KC);
        gs->row_lo.alloc(rp->n_rows); gs->row_hi.alloc(rp->n_rows); gs->krow.alloc(std::max<int64_t>(rp->n_pairs, 1));
        ISL_LAUNCH(h, k_gg_range_init, h->grid_for(rp->n_rows, 256), 256, 0, gs->row_lo.p, gs->row_hi.p, rp->n_rows);
        if (rp->n_pairs > 0)
            ISL_LAUNCH(h, k_gg_pos, h->grid_for(rp->n_pairs * gs->KC, 256), 256, 0, rp->pair.p, rp->n_pairs, rp->nr, ft.ds, gs->KC, compact, fc.ds, ncl,
                       ft.elem_eqn.p, fc.elem_eqn.p, h->rowptr.p, h->col.p, gs->pos.p, gs->row_lo.p, gs->row_hi.p, gs->krow.p, gs->KR, st.p);
        DevBuf<int> mw; mw.alloc(1);
        ISL_CUDA(cudaMemsetAsync(mw.p, 0, sizeof(int), h->stream));
        ISL_LAUNCH(h, k_gg_max_width, std::min(h->grid_for(rp->n_rows, 256), h->n_sm * 8), 256, 0, gs->row_lo.p, gs->row_hi.p, rp->n_rows, mw.p);
        int hst[2] = {0, 0};
        ISL_CUDA(cudaMemcpyAsync(hst, st.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaMemcpyAsync(&gs->max_width, mw.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        gs->max_len = hst[1];
        gs->ok = hst[0] == 0 && (size_t)(gs->max_len + 1) * sizeof(double) * 4 <= 200 * 1024;   // the row buffers of one warp (up to four rows) must fit a CTA
        if (getenv("ISL_VERBOSE"))
            fprintf(stderr, "[isl] atomic-free generic path (%d,%d,%s): %s, %lld (element, row) pairs, %d x %d per element, longest row %d, positions %.2f GB\n",
                    t, c, compact ? "one scalar per node pair" : "full block", gs->ok ? "ok" : "not eligible", (long long)rp->n_pairs, gs->KR, gs->KC,
                    gs->max_len, (double)rp->n_pairs * gs->KC * 2 / 1e9);
        if (getenv("ISL_VERBOSE")) fprintf(stderr, "[isl]   widest column range of the block inside a row: %d\n", gs->max_width);
    }
    GenGatherSet* out = gs->ok ? gs.get() : nullptr;
    h->gengathers[key] = std::move(gs);
    return out;
}

template <int G, int U>
void launch_gen_gather_t(isl_engine* h, const GenGatherParams& g) {
    // rows per CTA: as many sub-warps as fit 256 threads and 64 KB of row buffers
    int threads = 256;
    while (threads > 32 && (size_t)(threads / G) * g.buf_len * sizeof(double) > 64 * 1024) threads >>= 1;
    if (threads < G) threads = G;
    const int nsub = threads / G;
    const size_t smem = (size_t)nsub * g.buf_len * sizeof(double);
    ISL_CUDA(cudaFuncSetAttribute(k_gen_gather_rows<G, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    const int64_t nb = (g.n_rows + nsub - 1) / nsub;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nb, (int64_t)h->n_sm * 16));
    k_gen_gather_rows<G, U><<<grid, threads, smem, h->stream>>>(g);
    h->launches++;
    ISL_CUDA(cudaGetLastError());
}

void launch_gen_gather(isl_engine* h, GenGatherSet* gs, const AsmParams& a, int t, int c, bool store) {
    const FieldDev& ft = h->fields[t]; const FieldDev& fc = h->fields[c];
    RowPairs* rp = get_rowpairs(h, t);
    GenGatherParams g; std::memset(&g, 0, sizeof(g));
    g.pair = rp->pair.p; g.row_start = rp->row_start.p; g.pos = gs->pos.p; g.Kbuf = h->gen_kbuf.p;
    g.nr = rp->nr; g.dst = ft.ds; g.KR = gs->KR; g.KC = gs->KC; g.compact = gs->compact; g.nc = fc.ndpe; g.dsc = fc.ds; g.n_rows = rp->n_rows;
    g.rowptr = h->rowptr.p; g.val = h->val.p; g.rhs = h->rhs.p;
    g.ed_c = fc.elem_dof.p; g.st_c = fc.status.p; g.presc_c = fc.presc.p; g.val_c = fc.values.p; g.incremental = a.incremental;
    g.store = store ? 1 : 0;
    g.buf_len = (gs->max_len + 1) & ~1;
    if (h->gen_krow) g.krow = gs->krow.p;
    if (h->gen_range) { g.row_lo = gs->row_lo.p; g.row_hi = gs->row_hi.p; g.buf_len = (std::max(gs->max_width, 1) + 1) & ~1; }
    if (g.KC <= 8) launch_gen_gather_t<8, 1>(h, g);
    else if (g.KC <= 16) launch_gen_gather_t<16, 1>(h, g);
    else if (g.KC <= 32) launch_gen_gather_t<32, 1>(h, g);
    else launch_gen_gather_t<32, 3>(h, g);
}

void launch_hypel_sym(isl_engine* h, AsmParams& p) {
    const int nt = p.nt;
    // test nodes per register tile.  Few-node elements (P2 tetrahedra) get small tiles = more tiles = more threads per
    // element: their occupancy is bound by the shared memory per element, not by registers (ISL_HYPEL_MC overrides)
    int MC = (nt <= 10) ? h->hypel_mc_small : 6;
    if (MC != 2 && MC != 3 && MC != 5 && MC != 6) MC = (nt <= 10) ? 5 : 6;
    int ntiles = 0;
    for (int N = 0; N < nt; N++) ntiles += N / MC + 1;
    ISL_REQUIRE(ntiles <= 160, "element too large for the tile table");
    const HypelSymLayout L(p.npe, p.nq, nt);
    const size_t per = (size_t)L.per_elem * sizeof(double);
    ISL_REQUIRE(per <= 200 * 1024, "element too large for shared-memory staging");
    // one tile per thread: EB elements with EB * ntiles threads, CTA size = the multiple of 32 that holds the tiles.  The
    // variants with few registers (MC 2, 3 and the Q2 variant) take CTAs of at most 128 threads and the EB that keeps the
    // most tiles resident per SM (shared memory 220 KB, 64 K registers, 168 / 128 registers per thread)
    const bool small = (MC <= 3) || (MC == 6 && h->hypel_occ3 && ntiles <= 96);
    const int max_threads = small ? (MC == 6 ? 96 : 128) : 256;
    int EB = 1, best_score = -1;
    for (int eb = 1; eb * ntiles <= max_threads; eb++) {
        if ((size_t)eb * per > (small ? 110u : 100u) * 1024) break;
        const int threads = ((eb * ntiles + 31) / 32) * 32;
        const int by_smem = (int)((size_t)(220 * 1024) / ((size_t)eb * per + 1024));
        const int by_regs = 65536 / ((small ? (MC == 6 ? 168 : 128) : 232) * threads);
        const int score = std::min(std::min(by_smem, by_regs), 16) * eb * ntiles;
        if (score > best_score || (score == best_score && !small)) { best_score = score; EB = eb; }
    }
    const int best_nt = std::min(max_threads, ((EB * ntiles + 31) / 32) * 32);
    p.EB = EB;
    const size_t smem = per * EB;
    auto launch = [&](auto kernel) {
        ISL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ISL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        const int64_t nbatch = (h->n_owned + EB - 1) / EB;
        if (nbatch == 0) return;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nbatch, (int64_t)h->n_sm * 16));
        kernel<<<grid, best_nt, smem, h->stream>>>(p, ntiles);
        h->launches++;
        ISL_CUDA(cudaGetLastError());
    };
    if (getenv("ISL_VERBOSE")) fprintf(stderr, "[isl] hyperelastic tile kernel: MC %d, %d tiles per element, %d elements per CTA of %d threads, %.1f KB shared memory\n", MC, ntiles, EB, best_nt, smem / 1024.0);
    // Q2 hexahedra: one element per CTA of 96 threads, compiled for 168 registers: four CTAs = twelve warps per SM instead of
    // six (the kernel is bound by dependent-issue latency: 40.4 -> 33.5 ms at 64^3, profiles/r2/session36.log)
    if (MC == 6 && small) launch(k_tangent_hypel_sym<6, 96, 3>);
    else if (MC == 2) launch(k_tangent_hypel_sym<2, 128, 4>);
    else if (MC == 3) launch(k_tangent_hypel_sym<3, 128, 4>);
    else if (MC == 5) launch(k_tangent_hypel_sym<5>);
    else launch(k_tangent_hypel_sym<6>);
}

template <class K>
void launch_staged(isl_engine* h, K kernel, AsmParams& p, int tasks_per_elem = 0) {
    const size_t per = stage_doubles_per_elem(p, h->dim) * sizeof(double);
    const size_t budget = (size_t)h->stage_kb * 1024;   // shared memory per CTA: fewer elements per batch = more CTAs per SM
    ISL_REQUIRE(per <= 200 * 1024, "element too large for shared-memory staging");
    int EB = (int)std::max<size_t>(1, std::min<size_t>(budget / per, 32));
    if (tasks_per_elem > 0) {
        // strip tasks: the batch whose tasks fill whole passes of the 256 threads best (13 elements x 20 tasks = 260 would
        // run a second pass for four threads)
        int best = EB; double best_u = 0.;
        for (int b = EB; b >= std::max(1, EB / 2); b--) {
            const int tasks = b * tasks_per_elem, passes = (tasks + 255) / 256;
            const double u = (double)tasks / (passes * 256);
            if (u > best_u + 1e-9) { best_u = u; best = b; }
        }
        EB = best;
    }
    p.EB = EB;
    const size_t smem = per * EB;
    ISL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    const int64_t nbatch = (h->n_owned + EB - 1) / EB;
    if (nbatch == 0) return;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nbatch, (int64_t)h->n_sm * 8));
    kernel<<<grid, 256, smem, h->stream>>>(p);
    h->launches++;
    ISL_CUDA(cudaGetLastError());
}

void check_kernel_fields(isl_engine* h, int kid, int t, int c, bool tangent) {
    const FieldDev& ft = h->fields[t]; const FieldDev& fc = h->fields[c];
    switch (kid) {
        case ISL_K_MASS:
            ISL_REQUIRE(ft.ds == fc.ds, "Mass kernel: test and trial DoF sizes differ");
            ISL_REQUIRE(tangent, "base::kernel::Mass: only the matrix is implemented (no residual)");
            break;
        case ISL_K_LAPLACE:
        case ISL_K_VECTOR_LAPLACE:
            ISL_REQUIRE(ft.ds == fc.ds, "Laplace kernel: test and trial DoF sizes differ");
            if (!tangent && kid == ISL_K_LAPLACE) ISL_REQUIRE(fc.ds == 1, "heat::Laplace residual needs a scalar field");
            break;
        case ISL_K_HYPEL_STVENANT:
        case ISL_K_HYPEL_NEOHOOKE:
            ISL_REQUIRE(ft.ds == h->dim && fc.ds == h->dim, "HyperElastic kernel: DoF size must equal the space dimension");
            break;
        case ISL_K_PRESSURE_GRADIENT:
            ISL_REQUIRE(ft.ds == h->dim && fc.ds == 1, "PressureGradient: test = velocity (dim), trial = pressure (1)");
            break;
        case ISL_K_CONVECTION:
            ISL_REQUIRE(ft.ds == h->dim && fc.ds == h->dim, "fluid::Convection: test and trial must be velocity fields (DoF size = dimension)");
            break;
        case ISL_K_VELOCITY_DIVERGENCE:
            ISL_REQUIRE(ft.ds == 1 && fc.ds == h->dim, "VelocityDivergence: test = pressure (1), trial = velocity (dim)");
            break;
        default: throw IslError("unknown kernel id " + std::to_string(kid));
    }
}

void load_q1_tables(isl_engine* h) {
    if (h->q1_tables_loaded) return;
    const isl::Rule R = isl::make_rule(ISL_HEX, 3);
    const isl::Basis B(ISL_HEX, 1);
    std::vector<double> dN(8 * 8 * 3), N(8);
    std::vector<double> Nq(64);
    for (int q = 0; q < 8; q++) { B.eval(&R.p[q * 3], N.data(), &dN[q * 24]); for (int a = 0; a < 8; a++) Nq[q * 8 + a] = N[a]; }
    ISL_CUDA(cudaMemcpyToSymbolAsync(c_q1_dN, dN.data(), sizeof(double) * 192, 0, cudaMemcpyHostToDevice, h->stream));
    ISL_CUDA(cudaMemcpyToSymbolAsync(c_q1_N, Nq.data(), sizeof(double) * 64, 0, cudaMemcpyHostToDevice, h->stream));
    {
        std::vector<double> aff(6 * 36, 0.), nsum(8, 0.);
        const int al[6] = {0, 1, 2, 0, 0, 1}, be[6] = {0, 1, 2, 1, 2, 2};
        for (int c = 0; c < 6; c++)
            for (int a = 0; a < 8; a++)
                for (int b = a; b < 8; b++) {
                    double v = 0.;
                    for (int q = 0; q < 8; q++) {
                        v += dN[q * 24 + a * 3 + al[c]] * dN[q * 24 + b * 3 + be[c]];
                        if (al[c] != be[c]) v += dN[q * 24 + a * 3 + be[c]] * dN[q * 24 + b * 3 + al[c]];
                    }
                    aff[c * 36 + sym_idx(a, b)] = v;
                }
        for (int a = 0; a < 8; a++) for (int q = 0; q < 8; q++) nsum[a] += Nq[q * 8 + a];
        ISL_CUDA(cudaMemcpyToSymbolAsync(c_q1_aff, aff.data(), sizeof(double) * 216, 0, cudaMemcpyHostToDevice, h->stream));
        ISL_CUDA(cudaMemcpyToSymbolAsync(c_q1_Nsum, nsum.data(), sizeof(double) * 8, 0, cudaMemcpyHostToDevice, h->stream));
    }
    const double g[2] = {R.p[0], R.p[3]};  // 1-D abscissae in table order (point 0 = (g0,g0,g0), point 1 = (g1,g0,g0))
    const double n1[4] = {1. - g[0], g[0], 1. - g[1], g[1]};
    ISL_CUDA(cudaMemcpyToSymbolAsync(c_q1_n1, n1, sizeof(double) * 4, 0, cudaMemcpyHostToDevice, h->stream));
    ISL_CUDA(cudaMemcpyToSymbolAsync(c_q1_w, R.w.data(), sizeof(double) * 8, 0, cudaMemcpyHostToDevice, h->stream));
    ISL_CUDA(cudaStreamSynchronize(h->stream));
    h->q1_tables_loaded = true;
}

bool qualifies_q1(const isl_engine* h, int t, int c) {
    const FieldDev& ft = h->fields[t];
    return h->shape == ISL_HEX && h->geom_deg == 1 && t == c && ft.set && ft.deg == 1 && ft.ds == 1 && ft.dof_is_node &&
           !ft.has_masters;  // slaves of master DoFs take the generic kernels
}


// patch tables of the Q1 row kernel built on the device (isl_patch_dev.cuh); false: not eligible (caller falls back)
bool form_patches_device(isl_engine* h, FieldDev& f, PatchSet* ps) {
    const int64_t n = h->n_owned, nr = h->n_eqn, nn = h->n_nodes;
    if (n <= 0 || nr <= 0 || n * 8 >= ((int64_t)1 << 31) || h->dim != 3) return false;
    build_elem_eqn(h, f);
    cudaStream_t st = h->stream;
    // 1. row -> node, mean element extents, bounding box, numbering axis
    DevBuf<int32_t> row_node; row_node.alloc(nr);
    ISL_CUDA(cudaMemsetAsync(row_node.p, 0xff, nr * sizeof(int32_t), st));
    ISL_LAUNCH(h, k_pd_row_node, h->grid_for(nn, 256), 256, 0, f.eqn.p, nn, row_node.p);
    DevBuf<double> sums; sums.alloc(4);
    ISL_CUDA(cudaMemsetAsync(sums.p, 0, 4 * sizeof(double), st));
    const int64_t stride = std::max<int64_t>(1, n / 200000);
    ISL_LAUNCH(h, k_pd_extent, h->grid_for((n + stride - 1) / stride, 128), 128, 0, h->coords.p, h->conn.p, n, stride, sums.p);
    DevBuf<unsigned long long> bb; bb.alloc(9);
    ISL_CUDA(cudaMemsetAsync(bb.p, 0xff, 3 * sizeof(unsigned long long), st));
    ISL_CUDA(cudaMemsetAsync(bb.p + 3, 0, 6 * sizeof(unsigned long long), st));
    ISL_LAUNCH(h, k_bbox, std::min(h->grid_for(nn, 256), h->n_sm * 8), 256, 0, h->coords.p, nn, 3, bb.p, bb.p + 3);
    double hs[4];
    unsigned long long hb[9];
    ISL_CUDA(cudaMemcpyAsync(hs, sums.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
    ISL_CUDA(cudaStreamSynchronize(st));
    PdGrid g;
    for (int d = 0; d < 3; d++) g.h[d] = (hs[3] > 0 && hs[d] > 0) ? hs[d] / hs[3] : 1.0;
    DevBuf<double> dh; dh.alloc(3);
    ISL_CUDA(cudaMemcpyAsync(dh.p, g.h, 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    const int64_t vstride = std::max<int64_t>(1, nr / 100000);
    ISL_LAUNCH(h, k_pd_axis_votes, h->grid_for(nr / vstride + 1, 128), 128, 0, h->coords.p, row_node.p, nr, vstride, dh.p, bb.p + 6);
    ISL_CUDA(cudaMemcpyAsync(hb, bb.p, sizeof(hb), cudaMemcpyDeviceToHost, st));
    ISL_CUDA(cudaStreamSynchronize(st));
    auto ord2dbl = [](unsigned long long o) {
        const unsigned long long b = (o & 0x8000000000000000ull) ? (o & 0x7fffffffffffffffull) : ~o;
        double x; std::memcpy(&x, &b, sizeof(x)); return x;
    };
    int ax = 0;
    if (hb[7] > hb[6 + ax]) ax = 1;
    if (hb[8] > hb[6 + ax]) ax = 2;
    // 2. boxes: Bs x Bo x Bo lattice cells with Bs along the numbering axis (patch_stretch > 1) and about patch_rows rows
    const int R = std::max(16, h->patch_rows);
    // cubes in coordinates shrunk by patch_stretch along the numbering axis: Bs = stretch * Bo, Bs * Bo^2 ~ R
    const int Bo = std::max(1, (int)std::floor(std::cbrt((double)R / std::max(1.0, h->patch_stretch)) + 1e-9));
    const int Bs = std::max(1, R / (Bo * Bo));
    int64_t nb_total = 1;
    for (int d = 0; d < 3; d++) {
        g.lo[d] = ord2dbl(hb[d]);
        const double ext = ord2dbl(hb[3 + d]) - g.lo[d];
        const int64_t cells = (int64_t)std::floor(ext / g.h[d] + 0.5) + 1;
        g.B[d] = (d == ax) ? Bs : Bo;
        g.NB[d] = (int)std::max<int64_t>(1, (cells + g.B[d] - 1) / g.B[d]);
        nb_total *= g.NB[d];
    }
    if (nb_total >= ((int64_t)1 << 31)) return false;
    DevBuf<int32_t> key, key2, idx;
    key.alloc(nr); key2.alloc(nr); idx.alloc(nr); ps->rows.alloc(nr);
    ISL_LAUNCH(h, k_pd_row_keys, h->grid_for(nr, 256), 256, 0, h->coords.p, row_node.p, nr, g, key.p, idx.p);
    int bits = 1; while (((int64_t)1 << bits) < nb_total) bits++;
    size_t tb = 0;
    DevBuf<char> tmp;
    ISL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key2.p, idx.p, ps->rows.p, nr, 0, bits, st));
    tmp.alloc(tb);
    ISL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, key2.p, idx.p, ps->rows.p, nr, 0, bits, st));
    // run lengths of the sorted box keys = rows per patch
    DevBuf<int32_t> ukey, ucnt; DevBuf<int> nruns;
    ukey.alloc(nr); ucnt.alloc(nr); nruns.alloc(1);
    tb = 0;
    ISL_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tb, key2.p, ukey.p, ucnt.p, nruns.p, nr, st));
    if (tb > tmp.n) tmp.alloc(tb);
    ISL_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.p, tb, key2.p, ukey.p, ucnt.p, nruns.p, nr, st));
    int np = 0;
    ISL_CUDA(cudaMemcpyAsync(&np, nruns.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    ISL_CUDA(cudaStreamSynchronize(st));
    if (np <= 0) return false;
    ps->p_row_off.alloc(np + 1);
    ISL_CUDA(cudaMemsetAsync(ps->p_row_off.p, 0, sizeof(int32_t), st));
    tb = 0;
    ISL_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, ucnt.p, ps->p_row_off.p + 1, np, st));
    if (tb > tmp.n) tmp.alloc(tb);
    ISL_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, ucnt.p, ps->p_row_off.p + 1, np, st));
    h->launches += 8;
    key.release(); key2.release(); idx.release(); ukey.release(); ucnt.release();
    DevBuf<int32_t> patch_of_row, rank_of_row;
    patch_of_row.alloc(nr); rank_of_row.alloc(nr);
    ISL_LAUNCH(h, k_pd_patch_of_row, np, 64, 0, ps->p_row_off.p, np, ps->rows.p, patch_of_row.p, rank_of_row.p);
    // 3. element instances: unique (patch, element) pairs
    const int64_t nk = n * 8;
    DevBuf<uint64_t> k1, k2; DevBuf<int64_t> nsel;
    k1.alloc(nk); k2.alloc(nk); nsel.alloc(1);
    ISL_LAUNCH(h, k_pd_inst_keys, h->grid_for(nk, 256), 256, 0, f.elem_eqn.p, n, patch_of_row.p, k1.p);
    int pbits = 1; while (((int64_t)1 << pbits) < np + 1) pbits++;
    tb = 0;
    ISL_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, k1.p, k2.p, nk, 0, 64, st));
    if (tb > tmp.n) tmp.alloc(tb);
    ISL_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, k1.p, k2.p, nk, 0, 64, st));
    tb = 0;
    ISL_CUDA(cub::DeviceSelect::Unique(nullptr, tb, k2.p, k1.p, nsel.p, nk, st));
    if (tb > tmp.n) tmp.alloc(tb);
    ISL_CUDA(cub::DeviceSelect::Unique(tmp.p, tb, k2.p, k1.p, nsel.p, nk, st));
    int64_t n_inst = 0;
    ISL_CUDA(cudaMemcpyAsync(&n_inst, nsel.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    ISL_CUDA(cudaStreamSynchronize(st));
    {   // the invalid key (corners without an equation) sorts last
        uint64_t last = 0;
        if (n_inst > 0) ISL_CUDA(cudaMemcpy(&last, k1.p + n_inst - 1, sizeof(uint64_t), cudaMemcpyDeviceToHost));
        if (n_inst > 0 && last == ~0ull) n_inst--;
    }
    if (n_inst <= 0 || n_inst * 8 >= ((int64_t)1 << 31)) return false;
    h->launches += 4;
    k2.release();
    DevBuf<uint64_t> inst_keys; inst_keys.alloc(n_inst);
    ISL_CUDA(cudaMemcpyAsync(inst_keys.p, k1.p, n_inst * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    k1.release();
    ps->p_inst_off.alloc(np + 1);
    ISL_LAUNCH(h, k_pd_segment_offsets, (np + 1 + 127) / 128, 128, 0, inst_keys.p, n_inst, np, ps->p_inst_off.p);
    // 4. nodes of every patch: unique (patch, node) pairs of its instances
    const int64_t nnk = n_inst * 8;
    DevBuf<uint64_t> m1, m2; DevBuf<int32_t> inst_elem;
    m1.alloc(nnk); m2.alloc(nnk); inst_elem.alloc(n_inst);
    ISL_LAUNCH(h, k_pd_node_keys, h->grid_for(n_inst, 256), 256, 0, inst_keys.p, n_inst, h->conn.p, m1.p, inst_elem.p);
    tb = 0;
    ISL_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, m1.p, m2.p, nnk, 0, 64, st));
    if (tb > tmp.n) tmp.alloc(tb);
    ISL_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, m1.p, m2.p, nnk, 0, 64, st));
    tb = 0;
    ISL_CUDA(cub::DeviceSelect::Unique(nullptr, tb, m2.p, m1.p, nsel.p, nnk, st));
    if (tb > tmp.n) tmp.alloc(tb);
    ISL_CUDA(cub::DeviceSelect::Unique(tmp.p, tb, m2.p, m1.p, nsel.p, nnk, st));
    int64_t n_pn = 0;
    ISL_CUDA(cudaMemcpyAsync(&n_pn, nsel.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    ISL_CUDA(cudaStreamSynchronize(st));
    h->launches += 4;
    m2.release();
    ps->nodes.alloc(n_pn); ps->p_node_off.alloc(np + 1);
    ISL_LAUNCH(h, k_pd_nodes_from_keys, h->grid_for(n_pn, 256), 256, 0, m1.p, n_pn, ps->nodes.p);
    ISL_LAUNCH(h, k_pd_segment_offsets, (np + 1 + 127) / 128, 128, 0, m1.p, n_pn, np, ps->p_node_off.p);
    m1.release();
    // 5. local node indices, slot table
    ps->i_lnode.alloc((size_t)n_inst * 8);
    DevBuf<uint16_t> rslot; rslot.alloc((size_t)nr * 8);
    ISL_CUDA(cudaMemsetAsync(rslot.p, 0xff, (size_t)nr * 8 * sizeof(uint16_t), st));
    DevBuf<int> cnt; cnt.alloc(4);
    ISL_CUDA(cudaMemsetAsync(cnt.p, 0, 4 * sizeof(int), st));
    ISL_LAUNCH(h, k_pd_instances, h->grid_for(n_inst, 128), 128, 0, inst_keys.p, n_inst, h->conn.p, f.elem_eqn.p, ps->p_inst_off.p,
               ps->p_node_off.p, ps->p_row_off.p, ps->nodes.p, patch_of_row.p, rank_of_row.p, ps->i_lnode.p, rslot.p, cnt.p + 2);
    // sizes (offset arrays are small)
    std::vector<int32_t> ho(np + 1), hn(np + 1), hr(np + 1);
    ISL_CUDA(cudaMemcpyAsync(ho.data(), ps->p_inst_off.p, (np + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    ISL_CUDA(cudaMemcpyAsync(hn.data(), ps->p_node_off.p, (np + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    ISL_CUDA(cudaMemcpyAsync(hr.data(), ps->p_row_off.p, (np + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    int herr[2] = {0, 0};
    ISL_CUDA(cudaMemcpyAsync(herr, cnt.p + 2, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    ISL_CUDA(cudaStreamSynchronize(st));
    if (herr[0] || herr[1]) return false;
    int max_inst = 0, max_nodes = 0, max_rows = 0;
    for (int p = 0; p < np; p++) {
        max_inst = std::max(max_inst, ho[p + 1] - ho[p]); max_nodes = std::max(max_nodes, hn[p + 1] - hn[p]);
        max_rows = std::max(max_rows, hr[p + 1] - hr[p]);
    }
    const size_t sm = (size_t)7 * ((max_inst + 2) & ~1) * 8 + std::max((size_t)((max_nodes + 1) & ~1) * 24, (size_t)4 * RG_STAGE * 8);
    if (sm > (size_t)226 * 1024 || max_nodes >= 65535 || max_inst >= 65534) return false;
    // row tables (positions, eligibility) exactly as for host-formed patches
    ps->r_meta.alloc((size_t)nr * sizeof(RowMeta));
    ISL_LAUNCH(h, k_row_meta, np, 128, 0, 0, ps->p_row_off.p, ps->p_inst_off.p, ps->rows.p, rslot.p, inst_elem.p, h->conn.p, f.eqn.p,
               f.status.p, h->rowptr.p, h->col.p, reinterpret_cast<RowMeta*>(ps->r_meta.p), (int32_t*)nullptr, cnt.p, cnt.p + 1);
    int hc[2] = {0, 0};
    ISL_CUDA(cudaMemcpyAsync(hc, cnt.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    ISL_CUDA(cudaStreamSynchronize(st));
    if (hc[1]) return false;
    ps->lift_nodes.alloc((size_t)std::max(1, hc[0]) * 27);
    if (hc[0] > 0)
        ISL_LAUNCH(h, k_row_meta, np, 128, 0, 1, ps->p_row_off.p, ps->p_inst_off.p, ps->rows.p, rslot.p, inst_elem.p, h->conn.p, f.eqn.p,
                   f.status.p, h->rowptr.p, h->col.p, reinterpret_cast<RowMeta*>(ps->r_meta.p), ps->lift_nodes.p, cnt.p, cnt.p + 1);
    ISL_CUDA(cudaStreamSynchronize(st));
    ps->n_patches = np; ps->max_rows = max_rows; ps->max_nodes = max_nodes; ps->max_inst = max_inst; ps->max_entries = 0;
    ps->n_inst = n_inst; ps->n_elems = n; ps->redundancy = (double)n_inst / (double)n;
    ps->usable = true; ps->rows_ok = true;
    if (getenv("ISL_VERBOSE"))
        fprintf(stderr, "[isl] patches formed on the device: %d boxes of %d x %d x %d cells (long axis %d), max rows %d, max instances %d, "
                        "max nodes %d, element instances %.3fx, %d rows next to constrained nodes\n",
                np, g.B[0], g.B[1], g.B[2], ax, max_rows, max_inst, max_nodes, ps->redundancy, hc[0]);
    return true;
}

// build (or fetch) the patch decomposition of a scalar Q1-hex field; returns nullptr when the mesh does not fit
// the shared-memory accumulator (caller falls back to the atomic kernel)
PatchSet* get_patchset(isl_engine* h, int field) {
    auto it = h->patchsets.find(field);
    if (it != h->patchsets.end()) return it->second->usable ? it->second.get() : nullptr;
    auto ps = std::make_unique<PatchSet>();
    PatchSet* out = nullptr;
    FieldDev& f = h->fields[field];
    const int64_t n = h->n_owned;
    build_elem_eqn(h, f);
    if (h->q1_rows && !getenv("ISL_PATCH_HOST")) {
        // row kernel: everything on the device (isl_patch_dev.cuh); the host bisection below remains for the round-1 patch
        // kernels and as a fallback
        if (form_patches_device(h, f, ps.get())) {
            out = ps.get();
            h->patchsets[field] = std::move(ps);
            return out;
        }
        ps = std::make_unique<PatchSet>();
    }
    // 1. host copies: element equations / connectivity, CSR row pointer, row positions
    const int64_t nrow = h->n_eqn;
    std::vector<int32_t> heqn((size_t)n * 8), hconn((size_t)n * 8), hnode_eqn(h->n_nodes);
    std::vector<int64_t> hrowptr(h->n_eqn + 1);
    std::vector<double> hcoords((size_t)h->n_nodes * 3);
    ISL_CUDA(cudaMemcpyAsync(heqn.data(), f.elem_eqn.p, (size_t)n * 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ISL_CUDA(cudaMemcpyAsync(hconn.data(), h->conn.p, (size_t)n * 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ISL_CUDA(cudaMemcpyAsync(hrowptr.data(), h->rowptr.p, (h->n_eqn + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    ISL_CUDA(cudaMemcpyAsync(hnode_eqn.data(), f.eqn.p, h->n_nodes * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ISL_CUDA(cudaMemcpyAsync(hcoords.data(), h->coords.p, hcoords.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    ISL_CUDA(cudaStreamSynchronize(h->stream));
    // boxes must be compact in the mesh topology, not in physical space (anisotropic meshes): measure the mean element
    // extent along every axis and bisect in coordinates scaled by it
    double hmean[3] = {0., 0., 0.};
    {
        const int64_t stride = std::max<int64_t>(1, n / 200000);
        int64_t cnt = 0;
        for (int64_t e = 0; e < n; e += stride, cnt++)
            for (int d = 0; d < 3; d++) {
                double mn = 1e300, mx = -1e300;
                for (int a = 0; a < 8; a++) { const double v = hcoords[(size_t)hconn[(size_t)e * 8 + a] * 3 + d]; mn = std::min(mn, v); mx = std::max(mx, v); }
                hmean[d] += mx - mn;
            }
        for (int d = 0; d < 3; d++) hmean[d] = (cnt && hmean[d] > 0.) ? hmean[d] / (double)cnt : 1.0;
    }
    std::vector<double> rowxyz((size_t)nrow * 3, 0.);
    std::vector<int32_t> hperm; hperm.reserve(nrow);
    for (int64_t nd = 0; nd < h->n_nodes; nd++) {
        const int32_t r = hnode_eqn[nd];
        if (r < 0) continue;
        for (int d = 0; d < 3; d++) rowxyz[(size_t)r * 3 + d] = hcoords[(size_t)nd * 3 + d] / hmean[d];
        hperm.push_back(r);
    }
    if (h->patch_stretch != 1.0 && nrow > 1) {
        // longer runs of consecutive rows per box (longer contiguous pieces of the CSR arrays per patch): find the axis
        // along which consecutive equation numbers advance and shrink it before the bisection
        int64_t votes[3] = {0, 0, 0};
        const int64_t stride = std::max<int64_t>(1, nrow / 100000);
        for (int64_t r = 0; r + 1 < nrow; r += stride) {
            int best = -1, nz = 0;
            for (int d = 0; d < 3; d++) {
                const double dd = std::fabs(rowxyz[(size_t)(r + 1) * 3 + d] - rowxyz[(size_t)r * 3 + d]);
                if (dd > 0.25) { nz++; best = d; }
            }
            if (nz == 1) votes[best]++;
        }
        int ax = 0;
        if (votes[1] > votes[ax]) ax = 1;
        if (votes[2] > votes[ax]) ax = 2;
        if (votes[ax] > 0)
            for (int64_t r = 0; r < nrow; r++) rowxyz[(size_t)r * 3 + ax] /= h->patch_stretch;
    }
    hcoords.clear(); hcoords.shrink_to_fit();
    // shared memory per CTA: accumulator (27 entries per row on a hex lattice) + coordinates of the patch's nodes
    // (owned + halo) + row metadata.  R is shrunk until the estimate fits the budget; if the real patches still
    // overflow (irregular boxes) the partition is redone with a smaller R.
    const int smem_budget = (h->patch_ctas_per_sm >= 4 ? 55 : h->patch_ctas_per_sm == 3 ? 74 : h->patch_ctas_per_sm == 2 ? 112 : 224) * 1024 -
                            (h->patch_ws ? 2 * WS_STAGE_DOUBLES * 128 * 8 + 64 : 0);
    int rows_per_patch = h->patch_rows, cap_nodes = 0, cap_entries = 0;
    PatchHost P;
    for (int attempt = 0; attempt < 4; attempt++) {
        for (;; rows_per_patch -= 8) {
            if (h->q1_rows) { cap_nodes = 65534; cap_entries = 1 << 30; break; }   // the row kernels keep no accumulator
            const double c = std::cbrt((double)rows_per_patch) + 2.0;
            cap_nodes = (int)(c * c * c * 1.35) + 32;
            cap_entries = (smem_budget - cap_nodes * 24 - rows_per_patch * 24 - 256) / 8;
            if (cap_entries >= rows_per_patch * 27 || rows_per_patch <= 16) break;
        }
        // compact row boxes by recursive coordinate bisection, then the elements of every box (owner computes)
        std::vector<int32_t> perm_try = hperm;
        const int64_t n_leaves = std::max<int64_t>(1, ((int64_t)perm_try.size() + rows_per_patch - 1) / rows_per_patch);
        std::vector<int64_t> bounds(n_leaves + 1, 0);
        bounds[n_leaves] = (int64_t)perm_try.size();
        if (!perm_try.empty()) rcb_split(perm_try.data(), rowxyz.data(), 0, (int64_t)perm_try.size(), (int)n_leaves, 0, bounds.data(), 0);
        P = PatchHost();
        P.want_slots = h->q1_rows != 0;
        form_patches(perm_try, bounds, heqn, hconn, hrowptr, h->n_eqn, h->n_nodes, cap_entries, cap_nodes, P);
        if (h->q1_rows) {
            // row kernel: 7 doubles per element instance + max(coordinates, write-out staging); four CTAs per SM want <= 56 KB
            const size_t sm = (size_t)7 * ((P.max_inst + 2) & ~1) * 8 + std::max((size_t)((P.max_nodes + 1) & ~1) * 24, (size_t)4 * RG_STAGE * 8);
            if (!P.lattice || sm <= (size_t)56 * 1024 || rows_per_patch <= 32 || attempt == 3) break;
            rows_per_patch = (int)(rows_per_patch * 0.85);
            continue;
        }
        if (!P.lattice || (P.max_entries <= cap_entries && P.max_nodes <= cap_nodes) || rows_per_patch <= 16) break;
        rows_per_patch = (int)(rows_per_patch * 0.88);
    }
    rowxyz.clear(); rowxyz.shrink_to_fit();
    bool fits = P.lattice && P.max_entries <= cap_entries && P.max_nodes <= cap_nodes && P.max_nodes < 65535 && cap_entries > 0;
    if (h->q1_rows)
        fits = fits && P.max_inst < 65534 &&
               (size_t)7 * ((P.max_inst + 2) & ~1) * 8 + std::max((size_t)((P.max_nodes + 1) & ~1) * 24, (size_t)4 * RG_STAGE * 8) <= (size_t)226 * 1024;
    if (fits) {
        ps->n_patches = (int)P.inst_off.size() - 1;
        ps->max_entries = P.max_entries; ps->max_rows = P.max_rows; ps->max_nodes = P.max_nodes;
        ps->n_inst = (int64_t)P.inst_elem.size(); ps->n_elems = n;
        ps->redundancy = n ? (double)ps->n_inst / (double)n : 0.;
        upload_vec(h, ps->p_inst_off, P.inst_off); upload_vec(h, ps->p_row_off, P.row_off); upload_vec(h, ps->p_node_off, P.node_off);
        upload_vec(h, ps->rows, P.rows); upload_vec(h, ps->nodes, P.nodes);
        upload_vec(h, ps->i_lnode, P.lnode);
        DevBuf<int32_t> inst_elem; upload_vec(h, inst_elem, P.inst_elem);
        int err = 0;
        if (!h->q1_rows) {   // tables of the round-1 patch kernels only (64 position bytes per element instance)
            upload_vec(h, ps->soff, P.soff); upload_vec(h, ps->i_lrow, P.lrow);
            upload_vec(h, ps->p_run_off, P.run_off); upload_vec(h, ps->run_start, P.run_start); upload_vec(h, ps->run_soff, P.run_soff);
            ps->i_pos.alloc((size_t)ps->n_inst * 64);
            DevBuf<int> derr; derr.alloc(1);
            ISL_CUDA(cudaMemsetAsync(derr.p, 0, sizeof(int), h->stream));
            ISL_LAUNCH(h, k_inst_pos, h->grid_for(ps->n_inst * 64, 256), 256, 0, inst_elem.p, f.elem_eqn.p, ps->n_inst, h->rowptr.p, h->col.p,
                       ps->i_pos.p, derr.p);
            ISL_CUDA(cudaMemcpyAsync(&err, derr.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            ISL_CUDA(cudaStreamSynchronize(h->stream));
        }
        if (!err) { ps->usable = true; out = ps.get(); } else ps->n_patches = 0;
        if (!err && h->q1_rows && P.lattice && ps->n_patches > 0) {
            // row-gather tables (isl_rowgather.cuh): slots from the host, positions / eligibility on the device
            DevBuf<uint16_t> rslot; upload_vec(h, rslot, P.rslot);
            const size_t nr = P.rows.size();
            ps->r_meta.alloc(nr * sizeof(RowMeta));
            DevBuf<int> cnt; cnt.alloc(2);
            ISL_CUDA(cudaMemsetAsync(cnt.p, 0, 2 * sizeof(int), h->stream));
            ISL_LAUNCH(h, k_row_meta, ps->n_patches, 128, 0, 0, ps->p_row_off.p, ps->p_inst_off.p, ps->rows.p, rslot.p, inst_elem.p,
                       h->conn.p, f.eqn.p, f.status.p, h->rowptr.p, h->col.p, reinterpret_cast<RowMeta*>(ps->r_meta.p),
                       (int32_t*)nullptr, cnt.p, cnt.p + 1);
            int hc[2] = {0, 0};
            ISL_CUDA(cudaMemcpyAsync(hc, cnt.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            ISL_CUDA(cudaStreamSynchronize(h->stream));
            if (!hc[1]) {
                ps->lift_nodes.alloc((size_t)std::max(1, hc[0]) * 27);
                if (hc[0] > 0)
                    ISL_LAUNCH(h, k_row_meta, ps->n_patches, 128, 0, 1, ps->p_row_off.p, ps->p_inst_off.p, ps->rows.p, rslot.p, inst_elem.p,
                               h->conn.p, f.eqn.p, f.status.p, h->rowptr.p, h->col.p, reinterpret_cast<RowMeta*>(ps->r_meta.p),
                               ps->lift_nodes.p, cnt.p, cnt.p + 1);
                ISL_CUDA(cudaStreamSynchronize(h->stream));
                ps->rows_ok = true; ps->max_inst = P.max_inst;
            }
            if (getenv("ISL_VERBOSE")) fprintf(stderr, "[isl] row-gather tables: %s, %d rows next to constrained nodes\n", ps->rows_ok ? "ok" : "not eligible", hc[0]);
        }
    }
    if (getenv("ISL_VERBOSE"))
        fprintf(stderr, "[isl] patches: %d (<= %d rows), max entries %d (cap %d), max nodes %d (cap %d), element instances %.3fx, lattice %d%s\n",
                (int)P.inst_off.size() - 1, rows_per_patch, P.max_entries, cap_entries, P.max_nodes, cap_nodes,
                n ? (double)P.inst_elem.size() / (double)n : 0., (int)P.lattice, fits ? "" : " -> atomic fallback");
    {
        // CTA size of the all-affine kernel: the one whose element batches leave the fewest idle thread slots
        double best = -1.;
        for (int nt : {256, 288, 320}) {
            int64_t slots = 0;
            for (size_t q = 0; q + 1 < P.inst_off.size(); q++) slots += ((P.inst_off[q + 1] - P.inst_off[q] + nt - 1) / nt) * nt;
            const double u = slots ? (double)P.inst_elem.size() / (double)slots : 0.;
            if (u > best + 0.02) { best = u; ps->aff_nt = nt; }
        }
    }
    if (getenv("ISL_VERBOSE")) {
        // slot utilisation of the element batches for the CTA sizes the all-affine kernel is built for
        fprintf(stderr, "[isl] batch slot utilisation:");
        for (int nt : {128, 160, 192, 224, 256, 288, 320, 384, 512}) {
            int64_t slots = 0;
            for (size_t q = 0; q + 1 < P.inst_off.size(); q++) slots += ((P.inst_off[q + 1] - P.inst_off[q] + nt - 1) / nt) * nt;
            fprintf(stderr, " %d:%.3f", nt, slots ? (double)P.inst_elem.size() / (double)slots : 0.);
        }
        fprintf(stderr, "\n");
    }
    h->patchsets[field] = std::move(ps);
    return out;
}

// every owned element affine?  (exact test on the edge vectors, once per coordinate set)
void ensure_affine_state(isl_engine* h) {
    if (h->affine_state >= 0) return;
    DevBuf<int> flag; flag.alloc(1);
    ISL_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), h->stream));
    if (h->n_owned > 0) ISL_LAUNCH(h, k_check_affine, h->grid_for(h->n_owned, 256), 256, 0, h->coords.p, h->conn.p, h->n_owned, flag.p);
    int na = 0;
    ISL_CUDA(cudaMemcpyAsync(&na, flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    ISL_CUDA(cudaStreamSynchronize(h->stream));
    h->affine_state = na ? 0 : 1;
}

// tables of the two-kernel path for general Q1 elements (isl_rows_fromk.cuh), built on the device: locality-sorted
// element order (cub radix sort of the smallest equation number), row -> element table with the lattice check, per-row
// CSR positions.  ok = false: the mesh is not lattice-like (a row is local node a of two elements) -> generic kernels.
FromKSet* get_fromk(isl_engine* h, int field) {
    auto it = h->fromk_sets.find(field);
    if (it != h->fromk_sets.end()) return it->second->ok ? it->second.get() : nullptr;
    auto fk = std::make_unique<FromKSet>();
    FieldDev& f = h->fields[field];
    const int64_t n = h->n_owned, nr = h->n_eqn;
    build_elem_eqn(h, f);
    fk->n_rows = nr; fk->n_elems = n; fk->built = true;
    if (n > 0 && nr > 0 && n < ((int64_t)1 << 31)) {
        DevBuf<int32_t> key, key2, idx;
        key.alloc(n); key2.alloc(n); idx.alloc(n); fk->eorder.alloc(n);
        ISL_LAUNCH(h, k_fromk_elem_key, h->grid_for(n, 256), 256, 0, f.elem_eqn.p, n, key.p, idx.p);
        size_t tb = 0;
        ISL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, key.p, key2.p, idx.p, fk->eorder.p, n, 0, 32, h->stream));
        DevBuf<char> tmp; tmp.alloc(tb);
        ISL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, key.p, key2.p, idx.p, fk->eorder.p, n, 0, 32, h->stream));
        h->launches += 4;
        std::vector<int64_t> rb, eb;
        {   // chunks of the producer / consumer pipeline: ~pipe_rows rows each, element ranges by the sorted keys
            const int NC = (int)std::max<int64_t>(1, (nr + h->pipe_rows - 1) / h->pipe_rows);
            rb.resize(NC + 1);
            for (int c = 0; c <= NC; c++) rb[c] = std::min<int64_t>(nr, (int64_t)c * h->pipe_rows);
            upload_vec(h, fk->d_row_b, rb); fk->d_elem_b.alloc(NC + 1);
            ISL_LAUNCH(h, k_fromk_bounds, (NC + 1 + 127) / 128, 128, 0, key2.p, n, fk->d_row_b.p, NC + 1, fk->d_elem_b.p);
            eb.resize(NC + 1);
            ISL_CUDA(cudaMemcpyAsync(eb.data(), fk->d_elem_b.p, (NC + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
            ISL_CUDA(cudaStreamSynchronize(h->stream));
            fk->NC = NC;   // elements past eb[NC] touch no ACTIVE row
        }
        fk->row_pos.alloc((size_t)nr * 8);
        ISL_CUDA(cudaMemsetAsync(fk->row_pos.p, 0xff, (size_t)nr * 8 * sizeof(int32_t), h->stream));
        DevBuf<int> cnt; cnt.alloc(2);
        ISL_CUDA(cudaMemsetAsync(cnt.p, 0, 2 * sizeof(int), h->stream));
        ISL_LAUNCH(h, k_fromk_row_pos, h->grid_for(n, 256), 256, 0, fk->eorder.p, f.elem_eqn.p, n, fk->row_pos.p, cnt.p + 1);
        fk->meta.alloc((size_t)nr * sizeof(RowMeta));
        ISL_LAUNCH(h, k_fromk_row_meta, h->grid_for(nr, 128), 128, 0, 0, nr, fk->row_pos.p, fk->eorder.p, h->conn.p, f.eqn.p, f.status.p,
                   h->rowptr.p, h->col.p, reinterpret_cast<RowMeta*>(fk->meta.p), (int32_t*)nullptr, cnt.p, cnt.p + 1);
        int hc[2] = {0, 0};
        ISL_CUDA(cudaMemcpyAsync(hc, cnt.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        if (!hc[1]) {
            fk->lift_nodes.alloc((size_t)std::max(1, hc[0]) * 27);
            if (hc[0] > 0)
                ISL_LAUNCH(h, k_fromk_row_meta, h->grid_for(nr, 128), 128, 0, 1, nr, fk->row_pos.p, fk->eorder.p, h->conn.p, f.eqn.p,
                           f.status.p, h->rowptr.p, h->col.p, reinterpret_cast<RowMeta*>(fk->meta.p), fk->lift_nodes.p, cnt.p, cnt.p + 1);
            // pipeline tables: look-back of the rows, ring of K slots, queue segments E(0), E(1), R(0), E(2), R(1), ...
            const int NC = fk->NC;
            DevBuf<int> dback; dback.alloc(1);
            ISL_CUDA(cudaMemsetAsync(dback.p, 0, sizeof(int), h->stream));
            ISL_LAUNCH(h, k_fromk_maxback, h->grid_for(nr, 256), 256, 0, fk->row_pos.p, nr, fk->d_row_b.p, fk->d_elem_b.p, NC, dback.p);
            ISL_CUDA(cudaMemcpyAsync(&fk->maxback, dback.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            ISL_CUDA(cudaStreamSynchronize(h->stream));
            fk->slots = std::min(NC, std::max(4, fk->maxback + 4));
            if (h->fromk_pipeline && NC >= 2 && fk->maxback + 2 <= fk->slots) {
                int64_t cap = 0;
                std::vector<int32_t> et(NC), rt(NC), sc;
                std::vector<uint8_t> sk;
                std::vector<int64_t> sf;
                for (int c = 0; c < NC; c++) {
                    cap = std::max(cap, eb[c + 1] - eb[c]);
                    et[c] = (int32_t)((eb[c + 1] - eb[c] + 127) / 128); rt[c] = (int32_t)((rb[c + 1] - rb[c] + 127) / 128);
                }
                int64_t items = 0;
                auto push = [&](int kind, int c) { sf.push_back(items); sc.push_back(c); sk.push_back((uint8_t)kind); items += kind ? rt[c] : et[c]; };
                push(0, 0);
                for (int c = 1; c < NC; c++) { push(0, c); push(1, c - 1); }
                push(1, NC - 1);
                sf.push_back(items);
                fk->chunk_cap = (cap + 15) & ~(int64_t)15; fk->n_items = items; fk->n_seg = (int)sc.size();
                upload_vec(h, fk->seg_first, sf); upload_vec(h, fk->seg_chunk, sc); upload_vec(h, fk->seg_kind, sk);
                upload_vec(h, fk->e_tiles, et); upload_vec(h, fk->r_tiles, rt);
                fk->counters.alloc((size_t)4 + 2 * NC);
                fk->Kring.alloc((size_t)44 * fk->slots * fk->chunk_cap);
                fk->pipe_ok = true;
            } else {
                fk->K.alloc((size_t)44 * n);
            }
            ISL_CUDA(cudaStreamSynchronize(h->stream));
            fk->ok = true;
            if (h->fromk_tile_order) {   // launch order of the row tiles (see k_q1hex_rows_fromK)
                const int64_t n_tiles = (nr + 127) / 128;
                DevBuf<unsigned long long> bb; bb.alloc(6);
                ISL_CUDA(cudaMemsetAsync(bb.p, 0xff, 3 * sizeof(unsigned long long), h->stream));
                ISL_CUDA(cudaMemsetAsync(bb.p + 3, 0, 3 * sizeof(unsigned long long), h->stream));
                ISL_LAUNCH(h, k_bbox, std::min(h->grid_for(h->n_nodes, 256), h->n_sm * 8), 256, 0, h->coords.p, h->n_nodes, h->dim, bb.p, bb.p + 3);
                DevBuf<int32_t> tk, tk2, ti;
                tk.alloc(n_tiles); tk2.alloc(n_tiles); ti.alloc(n_tiles); fk->tile_perm.alloc(n_tiles);
                ISL_LAUNCH(h, k_fromk_tile_key, h->grid_for(n_tiles, 128), 128, 0, h->coords.p, h->conn.p, fk->eorder.p, fk->row_pos.p, nr, 128,
                           n_tiles, bb.p, bb.p + 3, tk.p, ti.p);
                size_t tbt = 0;
                ISL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tbt, tk.p, tk2.p, ti.p, fk->tile_perm.p, n_tiles, 0, 32, h->stream));
                DevBuf<char> tmpt; tmpt.alloc(tbt);
                ISL_CUDA(cub::DeviceRadixSort::SortPairs(tmpt.p, tbt, tk.p, tk2.p, ti.p, fk->tile_perm.p, n_tiles, 0, 32, h->stream));
                ISL_CUDA(cudaStreamSynchronize(h->stream));
            }
        }
        if (getenv("ISL_VERBOSE"))
            fprintf(stderr, "[isl] general Q1 path: %s, %d rows next to constrained nodes; %s (%d chunks, look-back %d, %d ring slots of %lld elements)\n",
                    fk->ok ? "ok" : "not eligible", hc[0], fk->pipe_ok ? "producer/consumer pipeline through L2" : "two kernels through HBM", fk->NC,
                    fk->maxback, fk->slots, (long long)fk->chunk_cap);
    }
    FromKSet* out = fk->ok ? fk.get() : nullptr;
    h->fromk_sets[field] = std::move(fk);
    return out;
}

PatchSet* get_patchset(isl_engine* h, int field);

// is one of the Q1 fast paths available for this field?  (builds its tables on first use)
bool q1_ready(isl_engine* h, int field) {
    if (!h->q1_rows) return !h->wide_slots && get_patchset(h, field) != nullptr;   // round-1 patch kernels: 32-bit positions
    ensure_affine_state(h);
    if (h->affine_state == 1) { PatchSet* ps = get_patchset(h, field); if (ps && ps->rows_ok) return true; }
    return get_fromk(h, field) != nullptr;   // also serves affine meshes whose patches are not eligible
}

template <bool MATRIX>
void launch_patch(isl_engine* h, PatchSet* ps, const FieldDev& ft, double factor, int incremental, int body, double f0);

// launch order of a patch set for the current exchange plan: interface patches first (built once per plan)
bool comm_patch_order(isl_engine* h, PatchSet* ps) {
    if (ps->perm_built) return ps->perm.p != nullptr;
    ps->perm_built = true;
    CommState& c = *h->comm;
    if (!c.iface_row.p || ps->n_patches == 0) return false;
    DevBuf<int32_t> flag; flag.alloc(ps->n_patches);
    ISL_LAUNCH(h, k_comm_patch_flags, ps->n_patches, 64, 0, ps->p_row_off.p, ps->rows.p, c.iface_row.p, ps->n_patches, flag.p);
    std::vector<int32_t> hf(ps->n_patches), perm; perm.reserve(ps->n_patches);
    ISL_CUDA(cudaMemcpyAsync(hf.data(), flag.p, hf.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    ISL_CUDA(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < ps->n_patches; k++) if (hf[k]) perm.push_back(k);
    ps->n_iface = (int)perm.size();
    for (int k = 0; k < ps->n_patches; k++) if (!hf[k]) perm.push_back(k);
    upload_vec(h, ps->perm, perm);
    if (getenv("ISL_VERBOSE")) fprintf(stderr, "[isl] exchange overlap: %d of %d patches own interface rows and are launched first\n", ps->n_iface, ps->n_patches);
    return true;
}

// one launch (affine row kernel) or two (general elements) of the Q1 hot path: stiffness + Dirichlet lift (matrix != 0)
// and / or body force
void launch_q1(isl_engine* h, int field, int matrix, double factor, int incremental, int body, double f0) {
    const FieldDev& ft = h->fields[field];
    if (!h->q1_rows) {
        PatchSet* ps = get_patchset(h, field);
        ISL_REQUIRE(ps, "internal: Q1 patch launch without a patch set");
        if (matrix) launch_patch<true>(h, ps, ft, factor, incremental, body, f0); else launch_patch<false>(h, ps, ft, 0., 0, body, f0);
        return;
    }
    ensure_affine_state(h);
    RowsParams q;
    std::memset(&q, 0, sizeof(q));
    q.coords = h->coords.p; q.status = ft.status.p; q.presc = ft.presc.p; q.values = ft.values.p; q.val = h->val.p; q.rhs = h->rhs.p;
    q.factor = factor; q.incremental = incremental; q.store_mode = h->val_is_zero ? 1 : 0; q.body = body; q.f0 = f0; q.matrix = matrix;
    PatchSet* ps = (h->affine_state == 1) ? get_patchset(h, field) : nullptr;
    if (ps && ps->rows_ok) {
        if (ps->n_patches == 0) return;
        q.p_inst_off = ps->p_inst_off.p; q.p_row_off = ps->p_row_off.p; q.p_node_off = ps->p_node_off.p;
        q.nodes = ps->nodes.p; q.i_lnode = ps->i_lnode.p;
        q.meta = reinterpret_cast<const RowMeta*>(ps->r_meta.p); q.lift_nodes = ps->lift_nodes.p;
        q.node_cap = (ps->max_nodes + 1) & ~1; q.inst_cap = (ps->max_inst + 2) & ~1; q.n_patches = ps->n_patches;
        const int nt = h->rows_threads == 256 ? 256 : 128;
        const size_t smem_r = (size_t)7 * q.inst_cap * 8 + std::max((size_t)q.node_cap * 24, (size_t)(nt / 32) * RG_STAGE * 8);
        const int per_sm = (int)std::min<size_t>(nt == 256 ? 2 : 4, (size_t)(227 * 1024) / (smem_r + 1024));
        q.resident = h->n_sm * std::max(1, per_sm);
        auto launch_range = [&](int base, int count, cudaStream_t st) {
            if (count <= 0) return;
            q.patch_base = base;
            if (nt == 256) {
                ISL_CUDA(cudaFuncSetAttribute(k_q1hex_rows_affine<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
                k_q1hex_rows_affine<256, 2><<<count, 256, smem_r, st>>>(q);
            } else {
                ISL_CUDA(cudaFuncSetAttribute(k_q1hex_rows_affine<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
                k_q1hex_rows_affine<128, 4><<<count, 128, smem_r, st>>>(q);
            }
            h->launches++;
            ISL_CUDA(cudaGetLastError());
        };
        // multi-GPU: the patches that own interface rows run on the (high-priority) communication stream, the exchange
        // is queued right behind them there, and the interior patches run on the engine stream at the same time; the
        // two sets write disjoint rows (isl_comm.cuh).  comm_join() brings the streams together again.
        if (h->comm && h->comm->plan && matrix && comm_patch_order(h, ps) && ps->n_iface > 0 && ps->n_iface < ps->n_patches) {
            CommState& c = *h->comm;
            q.perm = ps->perm.p;
            ISL_CUDA(cudaEventRecord(c.ev_ready, h->stream));
            ISL_CUDA(cudaStreamWaitEvent(c.stream, c.ev_ready, 0));
            launch_range(0, ps->n_iface, c.stream);
            c.iface_event_valid = true; c.join_pending = true;
            launch_range(ps->n_iface, ps->n_patches - ps->n_iface, h->stream);
        } else {
            launch_range(0, ps->n_patches, h->stream);
        }
        return;
    }
    FromKSet* fk = get_fromk(h, field);
    ISL_REQUIRE(fk, "internal: Q1 row launch without tables");
    FromKParams k;
    k.coords = h->coords.p; k.conn = h->conn.p; k.eorder = fk->eorder.p; k.n_elems = fk->n_elems;
    k.row_pos = fk->row_pos.p; k.meta = reinterpret_cast<const RowMeta*>(fk->meta.p); k.n_rows = fk->n_rows; k.K = fk->K.p;
    q.lift_nodes = fk->lift_nodes.p;
    k.r = q; k.matrix = matrix;
    k.tile_perm = fk->tile_perm.p;
    constexpr int NT = 128;
    const size_t smem_k = (size_t)(NT / 32) * RG_STAGE * 8;
    if (fk->pipe_ok) {
        // ONE persistent kernel: element tiles produce K into an L2-resident ring, row tiles consume it (isl_rows_fromk.cuh)
        PipeParams q2;
        k.K = fk->Kring.p;
        q2.k = k;
        q2.seg_first = fk->seg_first.p; q2.seg_chunk = fk->seg_chunk.p; q2.seg_kind = fk->seg_kind.p;
        q2.row_b = fk->d_row_b.p; q2.elem_b = fk->d_elem_b.p;
        q2.n_seg = fk->n_seg; q2.NC = fk->NC; q2.slots = fk->slots; q2.maxback = fk->maxback; q2.chunk_cap = fk->chunk_cap;
        ISL_CUDA(cudaMemsetAsync(fk->counters.p, 0, fk->counters.n * sizeof(int), h->stream));
        q2.next = reinterpret_cast<unsigned long long*>(fk->counters.p);
        q2.e_prefix = fk->counters.p + 2; q2.r_prefix = fk->counters.p + 3;
        q2.edone = fk->counters.p + 4; q2.rdone = fk->counters.p + 4 + fk->NC;
        q2.e_tiles = fk->e_tiles.p; q2.r_tiles = fk->r_tiles.p; q2.n_items = fk->n_items;
        ISL_CUDA(cudaFuncSetAttribute(k_q1hex_pipeline, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_k));
        ISL_LAUNCH(h, k_q1hex_pipeline, h->n_sm * 3, NT, smem_k, q2);
        return;
    }
    k.e_lo = 0; k.e_hi = fk->n_elems; k.r_lo = 0; k.r_hi = fk->n_rows;
    ISL_LAUNCH(h, k_q1hex_elemK, (unsigned)((fk->n_elems + 127) / 128), 128, 0, k);
    const unsigned rgrid = (unsigned)((fk->n_rows + NT - 1) / NT);
#define ISL_FROMK_ROWS(UU, MB)                                                                                          \
    do {                                                                                                               \
        ISL_CUDA(cudaFuncSetAttribute(k_q1hex_rows_fromK<NT, UU, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_k)); \
        ISL_LAUNCH(h, (k_q1hex_rows_fromK<NT, UU, MB>), rgrid, NT, smem_k, k);                                         \
    } while (0)
    switch (h->fromk_variant) {   // loads in flight per thread against warps per SM (ISL_FROMK_VARIANT, tuning)
        case 1: ISL_FROMK_ROWS(2, 6); break;
        case 2: ISL_FROMK_ROWS(1, 7); break;
        case 3: ISL_FROMK_ROWS(2, 5); break;
        case 5: ISL_FROMK_ROWS(4, 5); break;
        default: ISL_FROMK_ROWS(8, 4); break;   // 64 loads in flight per thread, 16 warps per SM: 3.65 ms (4/5: 3.78, 2/6: 4.04, 1/7: 4.51)
    }
#undef ISL_FROMK_ROWS
}

template <bool MATRIX>
void launch_patch(isl_engine* h, PatchSet* ps, const FieldDev& ft, double factor, int incremental, int body, double f0) {
    if (ps->n_patches == 0) return;  // no ACTIVE row: nothing to assemble
    PatchParams p;
    p.coords = h->coords.p;
    p.p_inst_off = ps->p_inst_off.p; p.p_row_off = ps->p_row_off.p; p.p_node_off = ps->p_node_off.p;
    p.rows = ps->rows.p; p.soff = ps->soff.p; p.nodes = ps->nodes.p;
    p.p_run_off = ps->p_run_off.p; p.run_start = ps->run_start.p; p.run_soff = ps->run_soff.p;
    p.i_lnode = ps->i_lnode.p; p.i_lrow = ps->i_lrow.p; p.i_pos = ps->i_pos.p; p.rowptr = h->rowptr.p;
    p.status = ft.status.p; p.presc = ft.presc.p; p.values = ft.values.p;
    p.val = h->val.p; p.rhs = h->rhs.p; p.factor = factor; p.incremental = incremental;
    p.store_mode = h->val_is_zero ? 1 : 0; p.n_patches = ps->n_patches;
    p.acc_cap = MATRIX ? ((ps->max_entries + 1) & ~1) : 0; p.row_cap = (ps->max_rows + 1) & ~1; p.node_cap = (ps->max_nodes + 1) & ~1;
    p.body = body; p.f0 = f0; p.fast = h->q1_fast; p.dbg = getenv("ISL_DBG") ? atoi(getenv("ISL_DBG")) : 0;
    p.prof = nullptr;
    DevBuf<unsigned long long>& profbuf = h->profbuf;   // per engine (ISL_PROF=1 cycle counters)
    const bool prof = MATRIX && getenv("ISL_PROF");
    if (prof) { profbuf.alloc(8); ISL_CUDA(cudaMemsetAsync(profbuf.p, 0, 64, h->stream)); p.prof = profbuf.p; }
    const size_t smem = (size_t)p.acc_cap * 8 + (size_t)p.node_cap * 24 + (size_t)p.row_cap * 16 + (size_t)(p.row_cap + 2) * 8 + 16;
    if (MATRIX && h->affine_kernel && (h->q1_fast & 2) && !h->patch_ws && h->shape == ISL_HEX) {
        ensure_affine_state(h);
        if (h->affine_state == 1 && h->affine_kernel) {
#define ISL_AFF_LAUNCH(NT, MINB)                                                                                        \
    do {                                                                                                               \
        ISL_CUDA(cudaFuncSetAttribute(k_q1hex_patch_affine<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        ISL_LAUNCH(h, (k_q1hex_patch_affine<NT, MINB>), ps->n_patches, NT, smem, p);                                   \
    } while (0)
#define ISL_AFF_LAUNCH_S(NT, MINB)                                                                                      \
    do {                                                                                                               \
        ISL_CUDA(cudaFuncSetAttribute(k_q1hex_patch_affine<NT, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        ISL_LAUNCH(h, (k_q1hex_patch_affine<NT, MINB, true>), ps->n_patches, NT, smem, p);                             \
    } while (0)
            const int nt = h->patch_threads_aff ? h->patch_threads_aff : ps->aff_nt;
            switch (h->patch_ctas_per_sm) {
                case 1: ISL_AFF_LAUNCH(512, 1); break;
                case 2:
                    if (nt == 192) ISL_AFF_LAUNCH(192, 2);
                    else if (nt == 224) ISL_AFF_LAUNCH(224, 2);
                    else if (nt == 288 && h->aff_split) ISL_AFF_LAUNCH_S(288, 2);
                    else if (nt == 320 && h->aff_split) ISL_AFF_LAUNCH_S(320, 2);
                    else if (nt == 288) ISL_AFF_LAUNCH(288, 2);
                    else if (nt == 320) ISL_AFF_LAUNCH(320, 2);
                    else if (nt == 384) ISL_AFF_LAUNCH(384, 2);
                    else if (h->aff_split) ISL_AFF_LAUNCH_S(256, 2);
                    else ISL_AFF_LAUNCH(256, 2);
                    break;
                case 3:
                    if (nt == 256) ISL_AFF_LAUNCH(256, 3);
                    else if (nt == 160) ISL_AFF_LAUNCH(160, 3);
                    else if (nt == 224) ISL_AFF_LAUNCH(224, 3);
                    else ISL_AFF_LAUNCH(192, 3);
                    break;
                default: ISL_AFF_LAUNCH(128, 4); break;
            }
#undef ISL_AFF_LAUNCH
#undef ISL_AFF_LAUNCH_S
            return;
        }
    }
    if (MATRIX && h->patch_ws) {
        const size_t smem_ws = (size_t)p.acc_cap * 8 + (size_t)p.node_cap * 24 + (size_t)p.row_cap * 16 +
                               (size_t)2 * WS_STAGE_DOUBLES * 128 * 8 + 4 * 8 + (size_t)(p.row_cap + 2) * 8 + 16;
        ISL_CUDA(cudaFuncSetAttribute(k_q1hex_patch_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws));
        ISL_LAUNCH(h, k_q1hex_patch_ws, ps->n_patches, 256, smem_ws, p);
    } else if (h->patch_threads == 128 && h->patch_ctas_per_sm >= 3) {
        ISL_CUDA(cudaFuncSetAttribute(k_q1hex_patch<128, MATRIX, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ISL_LAUNCH(h, (k_q1hex_patch<128, MATRIX, 3>), ps->n_patches, 128, smem, p);
    } else if (h->patch_threads == 128 && MATRIX && (h->q1_fast & 2)) {
        ISL_CUDA(cudaFuncSetAttribute(k_q1hex_patch<128, MATRIX, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ISL_LAUNCH(h, (k_q1hex_patch<128, MATRIX, 2, true>), ps->n_patches, 128, smem, p);
    } else if (h->patch_threads == 128) {
        ISL_CUDA(cudaFuncSetAttribute(k_q1hex_patch<128, MATRIX, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ISL_LAUNCH(h, (k_q1hex_patch<128, MATRIX, 2>), ps->n_patches, 128, smem, p);
    } else {
        ISL_CUDA(cudaFuncSetAttribute(k_q1hex_patch<256, MATRIX, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ISL_LAUNCH(h, (k_q1hex_patch<256, MATRIX, 1>), ps->n_patches, 256, smem, p);
    }
    if (prof) {
        unsigned long long t[8];
        ISL_CUDA(cudaMemcpyAsync(t, profbuf.p, 64, cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        const char* nm_std[8] = {"prologue", "zero", "load", "K", "lift", "scatter", "writeout", "total"};
        const char* nm_ws[8] = {"c:K+stage", "c:wait_empty", "s:load", "s:wait_full", "s:phases", "prologue", "writeout", "total"};
        const char* const* nm = h->patch_ws ? nm_ws : nm_std;
        fprintf(stderr, "[isl-prof] cycles per patch (thread 0):");
        for (int i = 0; i < 8; i++) fprintf(stderr, " %s %.0f", nm[i], (double)t[i] / ps->n_patches);
        fprintf(stderr, "\n");
    }
}

// the memset of the matrix values is postponed at isl_system_create; whoever accumulates into val materialises it
void materialize_zero(isl_engine* h) {
    if (h->val_zero_pending && h->nnz) ISL_CUDA(cudaMemsetAsync(h->val.p, 0, h->nnz * sizeof(double), h->stream));
    h->val_zero_pending = false;
}

// launch of the Q1 patch matrix kernel is deferred by one call so that an immediately following body force on the same
// field is fused into the same pass over the elements (the reference application order: stiffness, then body force)
// work queued on the communication stream (interface patches, exchange) becomes visible to the engine stream
void comm_join(isl_engine* h) {
    if (!h->comm || !h->comm->join_pending) return;
    ISL_CUDA(cudaEventRecord(h->comm->ev_done, h->comm->stream));
    ISL_CUDA(cudaStreamWaitEvent(h->stream, h->comm->ev_done, 0));
    h->comm->join_pending = false;
}

void flush_pending(isl_engine* h, int fuse_body, double f0) {
    comm_join(h);
    if (!h->pending_q1.active) return;
    h->pending_q1.active = false;
    const int t = h->pending_q1.field;
    const bool full_store = h->val_is_zero && h->pattern_pairs.size() == 1;
    if (full_store) h->val_zero_pending = false;  // every entry is written by a plain store
    else materialize_zero(h);
    launch_q1(h, t, 1, h->pending_q1.factor, h->pending_q1.incremental, fuse_body, f0);
    h->val_is_zero = false;
}

#include "isl_solver.cuh"
#include "isl_microbench.cuh"

}  // namespace

// =============================================================================
// C ABI
// =============================================================================
extern "C" {

const char* isl_last_error(void) { return g_error.c_str(); }
int isl_version(void) { return 100; }

int isl_engine_create(int device, isl_handle* out) {
    return guarded([&] {
        ISL_REQUIRE(out, "null output handle");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw IslError("no CUDA device available: the assembly engine has no CPU fallback");
        ISL_REQUIRE(device >= 0 && device < count, "device index out of range");
        ISL_CUDA(cudaSetDevice(device));
        auto h = std::make_unique<isl_engine>();
        h->device = device;
        ISL_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        cudaDeviceProp prop;
        ISL_CUDA(cudaGetDeviceProperties(&prop, device));
        h->n_sm = prop.multiProcessorCount;
        if (const char* m = getenv("ISL_Q1_MODE")) h->q1_mode = (std::string(m) == "atomic") ? 0 : 1;
        if (const char* m = getenv("ISL_Q1_FAST")) h->q1_fast = atoi(m);  // 0 reference order, 1 sum factorisation, 3 + affine shortcut
        if (const char* m = getenv("ISL_DEFER")) h->defer_launch = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_TANGENT_TILED")) h->tangent_tiled = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_TANGENT_SYM")) h->tangent_sym = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_ELEM_ORDER")) h->elem_order = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_HYPEL_GATHER")) h->hypel_gather = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_GEN_GATHER")) h->gen_gather = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_GEN_RANGE")) h->gen_range = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_GEN_TILE")) h->gen_tile = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_GEN_KROW")) h->gen_krow = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_GEN_SAMPLED")) h->gen_sampled = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_HYPEL_MC")) h->hypel_mc_small = atoi(m);
        if (const char* m = getenv("ISL_HYPEL_OCC3")) h->hypel_occ3 = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_BLOCK_SLOTS")) h->block_slots = std::max(0, std::min(2, atoi(m)));
        if (const char* m = getenv("ISL_FROMK_TILE_ORDER")) h->fromk_tile_order = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_SLOT64")) h->force_slot64 = atoi(m) != 0;
        if (const char* m = getenv("ISL_STAGE_KB")) h->stage_kb = std::max(8, std::min(200, atoi(m)));   // test knob: 64-bit slot maps at any size
        if (const char* m = getenv("ISL_AFFINE_KERNEL")) h->affine_kernel = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_AFF_THREADS")) h->patch_threads_aff = atoi(m);
        if (const char* m = getenv("ISL_AFF_SPLIT")) h->aff_split = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_PATCH_WS")) { h->patch_ws = atoi(m) ? 1 : 0; if (h->patch_ws) { h->patch_ctas_per_sm = 1; h->patch_threads = 256; if (!getenv("ISL_PATCH_ROWS")) h->patch_rows = 448; } }
        if (const char* m = getenv("ISL_Q1_ROWS")) h->q1_rows = atoi(m) ? 1 : 0;
        if (!h->q1_rows) { h->patch_rows = 400; h->patch_stretch = 1.0; }   // geometry of the round-1 patch kernels
        if (const char* m = getenv("ISL_ROWS_THREADS")) h->rows_threads = atoi(m);
        if (const char* m = getenv("ISL_FROMK_PIPELINE")) h->fromk_pipeline = atoi(m) ? 1 : 0;
        if (const char* m = getenv("ISL_PIPE_ROWS")) h->pipe_rows = std::max(1024, atoi(m));
        if (const char* m = getenv("ISL_FROMK_VARIANT")) h->fromk_variant = atoi(m);
        if (const char* m = getenv("ISL_PATCH_ROWS")) h->patch_rows = std::max(16, atoi(m));
        if (const char* m = getenv("ISL_PATCH_STRETCH")) h->patch_stretch = std::max(0.125, std::min(64.0, atof(m)));
        if (const char* m = getenv("ISL_PATCH_THREADS")) h->patch_threads = atoi(m) == 128 ? 128 : 256;
        if (const char* m = getenv("ISL_PATCH_CTAS")) h->patch_ctas_per_sm = std::max(1, std::min(4, atoi(m)));
        *out = h.release();
    });
}
int isl_engine_destroy(isl_handle h) {
    return guarded([&] {
        if (!h) return;
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        cudaStream_t s = h->stream;
        if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); cudaEventDestroy(h->ev_asm); cudaEventDestroy(h->ev_copy); }
        delete h;
        cudaStreamDestroy(s);
    });
}
int isl_engine_set_option(isl_handle h, const char* name, double value) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);
        const std::string n(name ? name : "");
        const int v = (int)value;
        if (n == "q1_rows") h->q1_rows = v ? 1 : 0;               // only useful when the engine was created with ISL_Q1_ROWS=1
        else if (n == "rows_threads") h->rows_threads = v;
        else if (n == "rows_ss") (void)v;   // retired variant; accepted for old sweep scripts
        else if (n == "affine_kernel") h->affine_kernel = v ? 1 : 0;
        else if (n == "aff_split") h->aff_split = v ? 1 : 0;
        else if (n == "aff_threads") h->patch_threads_aff = v;
        else if (n == "tangent_tiled") h->tangent_tiled = v ? 1 : 0;
        else if (n == "tangent_sym") h->tangent_sym = v ? 1 : 0;
        else if (n == "elem_order") h->elem_order = v ? 1 : 0;
        else if (n == "defer") h->defer_launch = v ? 1 : 0;
        else if (n == "gen_gather") h->gen_gather = v ? 1 : 0;
        else if (n == "gen_range") h->gen_range = v ? 1 : 0;
        else if (n == "gen_tile") h->gen_tile = v ? 1 : 0;
        else if (n == "gen_krow") h->gen_krow = v ? 1 : 0;
        else if (n == "gen_sampled") h->gen_sampled = v ? 1 : 0;
        else if (n == "hypel_gather") h->hypel_gather = v ? 1 : 0;
        else throw IslError("unknown option '" + n + "'");
    });
}
int isl_synchronize(isl_handle h) { return guarded([&] { ISL_CUDA(cudaSetDevice(h->device)); flush_pending(h); ISL_CUDA(cudaStreamSynchronize(h->stream)); }); }
int isl_flush(isl_handle h) { return guarded([&] { ISL_CUDA(cudaSetDevice(h->device)); flush_pending(h); }); }
void* isl_engine_stream(isl_handle h) { return (void*)h->stream; }
int64_t isl_kernel_launches(isl_handle h) { return h->launches; }

// ---- host tables ----
int isl_quadrature(int shape, int degree, double* w, double* p) {
    int n = -1;
    guarded([&] {
        const isl::Rule R = isl::make_rule(shape, degree);
        if (w) std::copy(R.w.begin(), R.w.end(), w);
        if (p) std::copy(R.p.begin(), R.p.end(), p);
        n = R.n;
    });
    return n;
}
int isl_shape_nfun(int shape, int degree) {
    int n = -1;
    guarded([&] { n = isl::Basis(shape, degree).nfun; });
    return n;
}
int isl_shape_eval(int shape, int degree, const double* xi, double* fun, double* grad) {
    return guarded([&] { isl::Basis(shape, degree).eval(xi, fun, grad); });
}
int isl_support_points(int shape, int degree, double* pts) {
    return guarded([&] { isl::Basis(shape, degree).support(pts); });
}

// ---- DoF handling ----
int isl_ndpe(int shape, int fe_deg) {
    int n = -1;
    guarded([&] { n = isl::fe_layout(shape, fe_deg).total; });
    return n;
}
int isl_dof_generate(int shape, int geom_deg, int64_t n_elems, const int32_t* conn, int fe_deg, int32_t* elem_dof,
                     int64_t* n_obj) {
    return guarded([&] { *n_obj = isl::dof_generate(shape, geom_deg, n_elems, conn, fe_deg, elem_dof); });
}
int isl_mesh_boundary(int shape, int geom_deg, int64_t n_elems, const int32_t* conn, int64_t* pairs, int64_t* n_pairs) {
    return guarded([&] {
        std::vector<int64_t> b;
        isl::mesh_boundary(shape, geom_deg, n_elems, conn, b);
        *n_pairs = (int64_t)b.size() / 2;
        if (pairs) std::copy(b.begin(), b.end(), pairs);
    });
}
int isl_boundary_dofs(int shape, int geom_deg, int dim, const double* coords, int64_t n_elems, const int32_t* conn,
                      int fe_deg, const int32_t* elem_dof, int64_t n_pairs, const int64_t* pairs, int32_t* obj,
                      double* x, int64_t* n) {
    return guarded([&] {
        (void)n_elems;
        std::vector<int32_t> o; std::vector<double> xx;
        isl::boundary_dofs(shape, geom_deg, dim, coords, conn, fe_deg, elem_dof, n_pairs, pairs, o, xx);
        *n = (int64_t)o.size();
        if (obj) { std::copy(o.begin(), o.end(), obj); std::copy(xx.begin(), xx.end(), x); }
    });
}
int isl_number_dofs(int64_t n_obj, int dof_size, const uint8_t* status, int64_t init, int64_t* eqn, int64_t* n_numbered) {
    return guarded([&] { *n_numbered = isl::number_dofs(n_obj, dof_size, status, init, eqn); });
}

// ---- mesh / fields ----
int isl_mesh_set(isl_handle h, int shape, int geom_deg, int dim, int64_t n_nodes, const double* coords, int64_t n_elems,
                 const int32_t* conn) {
    return guarded([&] {
        flush_pending(h);
        ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(dim == isl::shape_dim(shape), "only DIM == shape dimension is supported (no manifolds)");
        ISL_REQUIRE(dim == 2 || dim == 3, "dimension must be 2 or 3");
        const isl::Basis G(shape, geom_deg);
        h->shape = shape; h->geom_deg = geom_deg; h->dim = dim; h->npe = G.nfun;
        h->n_nodes = n_nodes; h->n_elems = n_elems; h->n_owned = n_elems; h->affine_state = -1;
        upload(h, h->coords, coords, (size_t)n_nodes * dim);
        upload(h, h->conn, conn, (size_t)n_elems * h->npe);
        for (auto& f : h->fields) f.reset();
        h->tables.clear();
        mark_stale_if_touched(h);
        invalidate_pattern(h);
        ISL_CUDA(cudaStreamSynchronize(h->stream));
    });
}
int isl_mesh_set_owned(isl_handle h, int64_t n_owned) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);
        ISL_REQUIRE(n_owned >= 0 && n_owned <= h->n_elems, "owned element count out of range");
        h->n_owned = n_owned; h->affine_state = -1;
        for (auto& f : h->fields) f.eorder.release();
        h->slotmaps.clear(); h->blockmaps.clear(); h->gathersets.clear(); h->rowpairs.clear(); h->gengathers.clear();
        h->patchsets.clear();
        h->fromk_sets.clear();
    });
}
int isl_mesh_update_coords(isl_handle h, const double* coords) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);
        ISL_REQUIRE(h->n_nodes > 0, "mesh not set");
        h->affine_state = -1;
        ISL_CUDA(cudaMemcpyAsync(h->coords.p, coords, (size_t)h->n_nodes * h->dim * sizeof(double), cudaMemcpyDefault, h->stream));
    });
}
int isl_field_set(isl_handle h, int field, int fe_deg, int dof_size, int64_t n_obj, const int32_t* elem_dof,
                  const int64_t* eqn, const uint8_t* status, const double* prescribed, const double* values) {
    return guarded([&] {
        flush_pending(h);
        ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(field >= 0 && field < 5, "field index out of range (0..4)");
        ISL_REQUIRE(h->n_elems > 0, "mesh must be set before fields");
        ISL_REQUIRE(dof_size >= 1 && dof_size <= 3, "dof_size must be 1..3");
        FieldDev& f = h->fields[field];
        f.reset();
        f.deg = fe_deg; f.ds = dof_size; f.n_obj = n_obj;
        f.ndpe = isl::Basis(h->shape, fe_deg).nfun;
        const size_t n = (size_t)n_obj * dof_size;
        upload(h, f.elem_dof, elem_dof, (size_t)h->n_elems * f.ndpe);
        // equation numbers are narrowed to int32 on the device
        std::vector<int64_t> e64(n);
        ISL_CUDA(cudaMemcpy(e64.data(), eqn, n * sizeof(int64_t), cudaMemcpyDefault));
        std::vector<int32_t> e32(n);
        for (size_t i = 0; i < n; i++) {
            ISL_REQUIRE(e64[i] < ((int64_t)1 << 31), "equation number exceeds int32");
            e32[i] = e64[i] < 0 ? -1 : (int32_t)e64[i];
        }
        upload_vec(h, f.eqn, e32);
        upload(h, f.status, status, n);
        upload(h, f.presc, prescribed, n);
        upload(h, f.values, values, n);
        f.set = true;
        if (f.ndpe == h->npe) {
            DevBuf<int> nd; nd.alloc(1);
            ISL_CUDA(cudaMemsetAsync(nd.p, 0, sizeof(int), h->stream));
            const int64_t cnt = h->n_elems * h->npe;
            ISL_LAUNCH(h, k_count_diff, h->grid_for(cnt, 256), 256, 0, f.elem_dof.p, h->conn.p, cnt, nd.p);
            int ndiff = 1;
            ISL_CUDA(cudaMemcpyAsync(&ndiff, nd.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            ISL_CUDA(cudaStreamSynchronize(h->stream));
            f.dof_is_node = (ndiff == 0);
        }
        mark_stale_if_touched(h);
        invalidate_pattern(h);
        // tables depend on the field degrees only, but drop those that refer to this field id
        for (auto it = h->tables.begin(); it != h->tables.end();)
            if (it->first[1] == field || it->first[2] == field) it = h->tables.erase(it); else ++it;
        ISL_CUDA(cudaStreamSynchronize(h->stream));
    });
}
int isl_field_set_constraints(isl_handle h, int field, int64_t n_con, const int64_t* con_dof, const int64_t* con_ptr,
                              const int64_t* master_eqn, const double* weight) {
    return guarded([&] {
        flush_pending(h);
        ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(field >= 0 && field < 5 && h->fields[field].set, "field not set");
        FieldDev& f = h->fields[field];
        f.reset_constraints();
        mark_stale_if_touched(h);
        invalidate_pattern(h);
        if (n_con <= 0) return;
        const size_t n = (size_t)f.n_obj * f.ds;
        // status and equation numbers on the host: slaves must be CONSTRAINED, masters must be equation numbers
        std::vector<uint8_t> st(n);
        ISL_CUDA(cudaMemcpyAsync(st.data(), f.status.p, n, cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        std::vector<int32_t> cnt(n + 1, 0);
        for (int64_t k = 0; k < n_con; k++) {
            ISL_REQUIRE(con_dof[k] >= 0 && (size_t)con_dof[k] < n, "constraint on a DoF component that does not exist");
            ISL_REQUIRE(st[con_dof[k]] == ISL_CONSTRAINED, "a DoF component with master DoFs must have status CONSTRAINED");
            ISL_REQUIRE(con_ptr[k + 1] >= con_ptr[k], "con_ptr must be non-decreasing");
            ISL_REQUIRE(cnt[con_dof[k] + 1] == 0, "DoF component constrained twice");
            cnt[con_dof[k] + 1] = (int32_t)(con_ptr[k + 1] - con_ptr[k]);
        }
        for (size_t k = 0; k < n; k++) cnt[k + 1] += cnt[k];
        std::vector<int32_t> cm((size_t)cnt[n]);
        std::vector<double> cw((size_t)cnt[n]);
        for (int64_t k = 0; k < n_con; k++)
            for (int64_t j = con_ptr[k]; j < con_ptr[k + 1]; j++) {
                ISL_REQUIRE(master_eqn[j] >= 0 && master_eqn[j] < ((int64_t)1 << 31), "master DoFs must be ACTIVE (equation number >= 0)");
                const size_t q = (size_t)cnt[con_dof[k]] + (size_t)(j - con_ptr[k]);
                cm[q] = (int32_t)master_eqn[j]; cw[q] = weight[j];
            }
        f.has_masters = !cm.empty();
        if (!f.has_masters) return;
        upload_vec(h, f.cptr, cnt); upload_vec(h, f.cmaster, cm); upload_vec(h, f.cweight, cw);
        f.h_cptr = cnt; f.h_cmaster = cm;
    });
}
int isl_field_update(isl_handle h, int field, const double* prescribed, const double* values) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);
        ISL_REQUIRE(field >= 0 && field < 5 && h->fields[field].set, "field not set");
        FieldDev& f = h->fields[field];
        const size_t bytes = (size_t)f.n_obj * f.ds * sizeof(double);
        if (prescribed) ISL_CUDA(cudaMemcpyAsync(f.presc.p, prescribed, bytes, cudaMemcpyDefault, h->stream));
        if (values) ISL_CUDA(cudaMemcpyAsync(f.values.p, values, bytes, cudaMemcpyDefault, h->stream));
    });
}

// ---- system ----
int isl_system_create(isl_handle h, int64_t n_eqn) {
    return guarded([&] {
        ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(n_eqn >= 0, "negative system size");
        h->pending_q1.active = false;  // a fresh solver discards what was queued for the old one
        if (n_eqn != h->n_eqn) { invalidate_pattern(h); h->n_eqn = n_eqn; h->rhs.alloc(n_eqn); }
        h->sys_pairs.clear();
        if (n_eqn) ISL_CUDA(cudaMemsetAsync(h->rhs.p, 0, n_eqn * sizeof(double), h->stream));
        h->val_zero_pending = true;   // memset of the matrix values postponed, see materialize_zero()
        h->val_is_zero = true;
        h->sys_stale = false; h->sys_released = false;
        if (h->ins_err.p) ISL_CUDA(cudaMemsetAsync(h->ins_err.p, 0, sizeof(int), h->stream));
    });
}
int isl_pattern_register(isl_handle h, int test_field, int trial_field) {
    return guarded([&] {
        flush_pending(h);
        ISL_CUDA(cudaSetDevice(h->device));
        ensure_pair(h, test_field, trial_field);
        if (qualifies_q1(h, test_field, trial_field) && h->q1_mode == 1 && q1_ready(h, test_field)) return;
        // (atomic-free generic path: its tables are built at the first assembly call, the slot map only if a kernel asks for it)
        if (h->gen_gather && !h->fields[test_field].has_masters && !h->fields[trial_field].has_masters) return;
        get_slotmap(h, test_field, trial_field);
    });
}

namespace {
// register strips of k_tangent (knob gen_tile): sets p.tile and returns the tasks per element, 0 = per-entry loop
int tile_tasks(isl_engine* h, AsmParams& p, int kid) {
    p.tile = 0;
    if (!h->gen_tile) return 0;
    if (kid == ISL_K_LAPLACE || kid == ISL_K_VECTOR_LAPLACE) { p.tile = 1; return p.nt * ((p.nc + 4) / 5); }
    if (kid == ISL_K_PRESSURE_GRADIENT || kid == ISL_K_VELOCITY_DIVERGENCE) { p.tile = 1; return p.nt * p.nc; }
    return 0;
}
void launch_tangent(isl_engine* h, AsmParams& p, int kid) {
    const int tpe = tile_tasks(h, p, kid);
    if (tpe > 0) { if (h->dim == 3) launch_staged(h, k_tangent_strips<3>, p, tpe); else launch_staged(h, k_tangent_strips<2>, p, tpe); }
    else { if (h->dim == 3) launch_staged(h, k_tangent<3>, p); else launch_staged(h, k_tangent<2>, p); }
}
// the third field of the tuple for kernels that read one (fluid::Convection): same basis as the trial field
void bind_aux(isl_engine* h, AsmParams& p, int kid, int c, int aux) {
    if (kid != ISL_K_CONVECTION) return;
    ISL_REQUIRE(aux >= 0 && aux < 5 && h->fields[aux].set, "fluid::Convection needs the advection velocity as third field of the tuple (isl_assemble_matrix_aux)");
    const FieldDev& fa = h->fields[aux]; const FieldDev& fc = h->fields[c];
    ISL_REQUIRE(fa.deg == fc.deg && fa.ndpe == fc.ndpe && fa.ds == h->dim, "fluid::Convection: the advection velocity must have the trial field's basis and DoF size = dimension");
    p.ed_a = fa.elem_dof.p; p.val_a = fa.values.p; p.dsa = fa.ds;
}
}  // namespace

int isl_assemble_matrix(isl_handle h, int kid, const double* params, int quad_deg, int t, int c, int incremental) {
    return isl_assemble_matrix_aux(h, kid, params, quad_deg, t, c, -1, incremental);
}
int isl_assemble_matrix_aux(isl_handle h, int kid, const double* params, int quad_deg, int t, int c, int aux, int incremental) {
    return guarded([&] {
        ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(h->n_eqn >= 0, "isl_system_create must be called first");
        ISL_REQUIRE(t >= 0 && t < 5 && c >= 0 && c < 5, "field index out of range");
        require_live_system(h);
        flush_pending(h);
        ensure_pair(h, t, c);
        check_kernel_fields(h, kid, t, c, true);
        const FieldDev& ft = h->fields[t];
        // hot path: Q1 hex, scalar Laplace, 2x2x2 rule
        if ((kid == ISL_K_LAPLACE || kid == ISL_K_VECTOR_LAPLACE) && qualifies_q1(h, t, c) && (quad_deg == 2 || quad_deg == 3)) {
            load_q1_tables(h);
            const double factor = params ? params[0] : 1.0;
            if (h->q1_mode == 1) {
                if (q1_ready(h, t)) {
                    h->pending_q1.active = true; h->pending_q1.field = t; h->pending_q1.factor = factor;
                    h->pending_q1.incremental = incremental;
                    if (!h->defer_launch) flush_pending(h);
                    return;
                }
            }
            if (!h->wide_slots) {   // (the one-thread-per-element kernel reads 32-bit slots; larger systems take k_tangent)
                materialize_zero(h);
                h->val_is_zero = false;
                Q1Params q;
                q.coords = h->coords.p; q.conn = h->conn.p; q.n_elems = h->n_owned; q.slot = get_slotmap(h, t, c)->s32.p;
                q.eqn = ft.eqn.p; q.status = ft.status.p; q.presc = ft.presc.p; q.values = ft.values.p;
                q.val = h->val.p; q.rhs = h->rhs.p; q.factor = params ? params[0] : 1.0; q.incremental = incremental;
                const int block = 128;
                const int64_t grid = (h->n_owned + block - 1) / block;
                if (grid > 0) ISL_LAUNCH(h, k_q1hex_laplace, (unsigned)grid, block, 0, q);
                return;
            }
        }
        AsmParams p; std::memset(&p, 0, sizeof(p));
        fill_common(h, p, quad_deg, t, c);
        const bool hypel = (kid == ISL_K_HYPEL_STVENANT || kid == ISL_K_HYPEL_NEOHOOKE);
        const bool sym_kernel = h->tangent_sym && hypel && ft.ds == 3 && h->dim == 3 && t == c && !ft.has_masters && ft.ndpe <= 27;
        if (sym_kernel && h->hypel_gather) {
            // atomic-free: element matrices to memory, CSR rows gathered by one warp each (isl_gather.cuh)
            if (GatherSet* gs = get_gatherset(h, t)) {
                const bool store = h->val_is_zero && h->pattern_pairs.size() == 1;   // nothing in the rows yet, nothing else will come
                if (store) h->val_zero_pending = false; else materialize_zero(h);
                h->val_is_zero = false;
                p.kernel_id = kid; p.incremental = incremental;
                p.p0 = params ? params[0] : 0.; p.p1 = params ? params[1] : 0.;
                p.need_gt = 1; p.need_gc = 1; p.nqdata = 81;
                p.kout = gs->Kbuf.p;
                launch_hypel_sym(h, p);
                launch_gather_rows(h, gs, p, ft, store);
                return;
            }
        }
        const bool gen_kid = gen_gather_compact(kid) || kid == ISL_K_PRESSURE_GRADIENT || kid == ISL_K_VELOCITY_DIVERGENCE;
        if (h->gen_gather && gen_kid) {
            // atomic-free: k_tangent stores the element matrices, a sub-warp per CSR row gathers them (isl_gather.cuh)
            if (GenGatherSet* gs = get_gengather(h, t, c, gen_gather_compact(kid) ? 1 : 0)) {
                const bool store = h->val_is_zero;   // fresh system: every row is written completely, zeros where nothing contributes
                if (store) h->val_zero_pending = false; else materialize_zero(h);
                h->val_is_zero = false;
                const size_t need = (size_t)h->n_owned * gs->KR * gs->KC;
                if (h->gen_kbuf.n < need) h->gen_kbuf.alloc(need);   // (stream-ordered free: the previous gather has been queued before)
                p.kernel_id = kid; p.incremental = incremental;
                p.p0 = params ? params[0] : 0.;
                p.need_gt = (kid != ISL_K_VELOCITY_DIVERGENCE && kid != ISL_K_MASS);
                p.need_gc = (kid == ISL_K_VELOCITY_DIVERGENCE) || (kid != ISL_K_PRESSURE_GRADIENT && kid != ISL_K_MASS);
                p.nqdata = (kid == ISL_K_CONVECTION ? 4 : 0);
                bind_aux(h, p, kid, c, aux);
                p.kout = h->gen_kbuf.p;
                launch_tangent(h, p, kid);
                launch_gen_gather(h, gs, p, t, c, store);
                return;
            }
        }
        materialize_zero(h);
        h->val_is_zero = false;
        // one position per node pair instead of a slot per entry (vector fields without slaves of master DoFs); the
        // per-entry path of irregular pairs (constrained nodes) then searches the row
        const bool tiled = h->tangent_tiled && hypel && ft.ds == h->dim && !sym_kernel;   // (experimental kernel: per-entry slots)
        // (measured, session29: the generic kernels gain -- Stokes UU block 15.2 -> 13.5 ms --, the hyperelastic tile kernel does
        // not -- C3 49.6 / 49.2 ms, C4 61.0 / 62.8 ms --, so it keeps its per-entry slots unless ISL_BLOCK_SLOTS=2)
        if (tiled || (sym_kernel && h->block_slots < 2) || !bind_blockmap(h, p, t, c)) bind_slots(h, p, t, c);
        p.kernel_id = kid; p.incremental = incremental;
        p.p0 = params ? params[0] : 0.; p.p1 = (params && (kid == ISL_K_HYPEL_STVENANT || kid == ISL_K_HYPEL_NEOHOOKE)) ? params[1] : 0.;
        p.need_gt = (kid != ISL_K_VELOCITY_DIVERGENCE && kid != ISL_K_MASS);
        p.need_gc = (kid == ISL_K_VELOCITY_DIVERGENCE) || (kid != ISL_K_PRESSURE_GRADIENT && kid != ISL_K_MASS);
        p.nqdata = (kid == ISL_K_HYPEL_STVENANT || kid == ISL_K_HYPEL_NEOHOOKE) ? 81 : (kid == ISL_K_CONVECTION ? 4 : 0);
        bind_aux(h, p, kid, c, aux);
        if (sym_kernel) {
            launch_hypel_sym(h, p);
            return;
        }
        if (h->tangent_tiled && (kid == ISL_K_HYPEL_STVENANT || kid == ISL_K_HYPEL_NEOHOOKE) && ft.ds == h->dim) {
            if (h->dim == 3) launch_staged(h, k_tangent_hypel_tiled<3>, p); else launch_staged(h, k_tangent_hypel_tiled<2>, p);
            return;
        }
        launch_tangent(h, p, kid);
    });
}

int isl_assemble_matrix_sampled(isl_handle h, int kid, const double* values, int quad_deg, int t, int c, int incremental) {
    return guarded([&] {
        ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(h->n_eqn >= 0, "isl_system_create must be called first");
        ISL_REQUIRE(t >= 0 && t < 5 && c >= 0 && c < 5, "field index out of range");
        ISL_REQUIRE(kid == ISL_K_LAPLACE || kid == ISL_K_VECTOR_LAPLACE, "sampled material factors: Laplace kernels only");
        ISL_REQUIRE(values != nullptr, "no conductivity values");
        require_live_system(h);
        flush_pending(h);
        ensure_pair(h, t, c);
        check_kernel_fields(h, kid, t, c, true);
        AsmParams p; std::memset(&p, 0, sizeof(p));
        fill_common(h, p, quad_deg, t, c);
        DevBuf<double> kq;   // host or device pointer
        upload(h, kq, values, (size_t)h->n_owned * p.nq);
        p.kernel_id = kid; p.incremental = incremental; p.kq = kq.p;
        p.need_gt = 1; p.need_gc = 1; p.nqdata = 0;
        GenGatherSet* gs = (h->gen_gather && h->gen_sampled) ? get_gengather(h, t, c, 1) : nullptr;
        if (gs) {   // atomic-free, as in isl_assemble_matrix_aux
            const bool store = h->val_is_zero;
            if (store) h->val_zero_pending = false; else materialize_zero(h);
            h->val_is_zero = false;
            const size_t need = (size_t)h->n_owned * gs->KR * gs->KC;
            if (h->gen_kbuf.n < need) h->gen_kbuf.alloc(need);
            p.kout = h->gen_kbuf.p;
            launch_tangent(h, p, kid);
            launch_gen_gather(h, gs, p, t, c, store);
        } else {
            materialize_zero(h);
            h->val_is_zero = false;
            bind_slots(h, p, t, c);
            launch_tangent(h, p, kid);
        }
        ISL_CUDA(cudaStreamSynchronize(h->stream));   // kq is released when this function returns
    });
}

int isl_assemble_residual(isl_handle h, int kid, const double* params, int quad_deg, int t, int c, double factor) {
    return isl_assemble_residual_aux(h, kid, params, quad_deg, t, c, -1, factor);
}
int isl_assemble_residual_aux(isl_handle h, int kid, const double* params, int quad_deg, int t, int c, int aux, double factor) {
    return guarded([&] {
        ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(h->n_eqn >= 0, "isl_system_create must be called first");
        ISL_REQUIRE(t >= 0 && t < 5 && c >= 0 && c < 5, "field index out of range");
        require_live_system(h);
        flush_pending(h);
        check_kernel_fields(h, kid, t, c, false);
        AsmParams p; std::memset(&p, 0, sizeof(p));
        fill_common(h, p, quad_deg, t, c);
        p.kernel_id = kid; p.factor = factor;
        p.p0 = params ? params[0] : 0.; p.p1 = (params && (kid == ISL_K_HYPEL_STVENANT || kid == ISL_K_HYPEL_NEOHOOKE)) ? params[1] : 0.;
        p.need_gt = (kid != ISL_K_VELOCITY_DIVERGENCE);
        p.need_gc = (kid != ISL_K_PRESSURE_GRADIENT);
        p.nqdata = 9;
        bind_aux(h, p, kid, c, aux);
        if (h->dim == 3) launch_staged(h, k_force<3>, p); else launch_staged(h, k_force<2>, p);
    });
}

int isl_assemble_bodyforce(isl_handle h, const double* f, int quad_deg, int t) {
    return guarded([&] {
        ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(h->n_eqn >= 0, "isl_system_create must be called first");
        ISL_REQUIRE(t >= 0 && t < 5 && h->fields[t].set, "field not set");
        require_live_system(h);
        {
            const FieldDev& ft = h->fields[t];
            if (h->q1_mode == 1 && h->shape == ISL_HEX && h->geom_deg == 1 && ft.deg == 1 && ft.ds == 1 && ft.dof_is_node &&
                !ft.has_masters && (quad_deg == 2 || quad_deg == 3) && h->pattern_pairs.count({t, t})) {
                load_q1_tables(h);
                if (h->pending_q1.active && h->pending_q1.field == t) { flush_pending(h, 1, f[0]); return; }
                flush_pending(h);
                if (q1_ready(h, t)) { launch_q1(h, t, 0, 0., 0, 1, f[0]); return; }
            }
        }
        flush_pending(h);
        AsmParams p; std::memset(&p, 0, sizeof(p));
        fill_common(h, p, quad_deg, t, t);
        p.body = 1; p.factor = 1.0;
        for (int d = 0; d < h->fields[t].ds; d++) p.f[d] = f[d];
        p.need_gt = 0; p.need_gc = 0; p.nqdata = 0;
        if (h->dim == 3) launch_staged(h, k_force<3>, p); else launch_staged(h, k_force<2>, p);
    });
}

int isl_assemble_bodyforce_sampled(isl_handle h, const double* values, int quad_deg, int t) {
    return guarded([&] {
        ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(h->n_eqn >= 0, "isl_system_create must be called first");
        ISL_REQUIRE(t >= 0 && t < 5 && h->fields[t].set, "field not set");
        ISL_REQUIRE(values != nullptr, "no force values");
        require_live_system(h);
        flush_pending(h);
        AsmParams p; std::memset(&p, 0, sizeof(p));
        fill_common(h, p, quad_deg, t, t);
        DevBuf<double> fq;   // host or device pointer
        upload(h, fq, values, (size_t)h->n_owned * p.nq * h->fields[t].ds);
        p.body = 2; p.factor = 1.0; p.fq = fq.p;
        p.need_gt = 0; p.need_gc = 0; p.nqdata = 0;
        if (h->dim == 3) launch_staged(h, k_force<3>, p); else launch_staged(h, k_force<2>, p);
        ISL_CUDA(cudaStreamSynchronize(h->stream));   // fq is released when this function returns
    });
}

// ---- DoF-object ids on the device (isl_dof_dev.cuh) ----
namespace {
int64_t dof_generate_device(isl_engine* h, int fe_deg, int32_t* d_elem_dof) {
    const int shape = h->shape, dim = h->dim, npe = h->npe;
    const int64_t ne = h->n_elems;
    const isl::FELayout L = isl::fe_layout(shape, fe_deg);
    if (fe_deg == h->geom_deg) {   // isoparametric: DoF id = node id
        DevBuf<int> mx; mx.alloc(1);
        ISL_CUDA(cudaMemsetAsync(mx.p, 0xff, sizeof(int), h->stream));
        ISL_LAUNCH(h, k_dg_copy_conn, h->grid_for(ne * npe, 256), 256, 0, h->conn.p, ne * npe, d_elem_dof, mx.p);
        int m = -1;
        ISL_CUDA(cudaMemcpyAsync(&m, mx.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        return (int64_t)m + 1;
    }
    int64_t next = 0;
    const auto t_start = std::chrono::steady_clock::now();
    for (int nf = 0; nf <= dim; nf++) {
        const int stride = L.per[nf], nfaces = L.count[nf];
        if (stride == 0 || nfaces == 0) continue;
        if (nf == dim) {   // interior DoFs, private to the element
            ISL_LAUNCH(h, k_dg_interior, h->grid_for(ne * stride * nfaces, 256), 256, 0, ne, stride * nfaces, L.total, L.begin[nf], next, d_elem_dof);
            next += ne * stride * nfaces;
            continue;
        }
        const int64_t n = ne * nfaces;
        ISL_REQUIRE(n < ((int64_t)1 << 32), "too many n-faces for 32-bit item numbers: partition the mesh first");
        DgTopo T; std::memset(&T, 0, sizeof(T));
        T.nfaces = nfaces; T.nv = std::min(4, isl::nface_num_vertices(shape, nf));
        for (int f = 0; f < nfaces; f++) for (int j = 0; j < T.nv; j++) T.vert[f][j] = isl::nface_vertex(shape, nf, f, j);
        DevBuf<uint64_t> khi, klo, k2, k3; DevBuf<uint32_t> idx, idx2, start, first, fresh, rank;
        khi.alloc(n); klo.alloc(n); k2.alloc(n); idx.alloc(n); idx2.alloc(n);
        ISL_LAUNCH(h, k_dg_keys, h->grid_for(n, 256), 256, 0, h->conn.p, ne, npe, T, khi.p, klo.p, idx.p);
        size_t tb = 0, tb2 = 0;
        ISL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, khi.p, k2.p, idx.p, idx2.p, n, 0, 64, h->stream));
        DevBuf<char> tmp; tmp.alloc(tb);
        const uint32_t* sorted_idx;
        if (T.nv > 2) {   // 128-bit key: least significant half first, both sorts stable
            ISL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, klo.p, k2.p, idx.p, idx2.p, n, 0, 64, h->stream));
            k3.alloc(n);
            ISL_LAUNCH(h, k_dg_gather64, h->grid_for(n, 256), 256, 0, khi.p, idx2.p, n, k3.p);
            ISL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, k3.p, k2.p, idx2.p, idx.p, n, 0, 64, h->stream));
            sorted_idx = idx.p;
        } else {
            ISL_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, khi.p, k2.p, idx.p, idx2.p, n, 0, 64, h->stream));
            sorted_idx = idx2.p;
        }
        h->launches += 6;
        // k2 = sorted most significant halves; run heads, run start per member, first visitor per item
        start.alloc(n); first.alloc(n); fresh.alloc(n); rank.alloc(n);
        ISL_LAUNCH(h, k_dg_heads, h->grid_for(n, 256), 256, 0, k2.p, klo.p, sorted_idx, n, start.p);
        ISL_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb2, start.p, start.p, DgMax(), n, h->stream));
        if (tb2 > tmp.n) tmp.alloc(tb2);
        ISL_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, tb2, start.p, start.p, DgMax(), n, h->stream));
        ISL_LAUNCH(h, k_dg_first, h->grid_for(n, 256), 256, 0, sorted_idx, start.p, n, first.p, fresh.p);
        ISL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, fresh.p, rank.p, n, h->stream));
        if (tb2 > tmp.n) tmp.alloc(tb2);
        ISL_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb2, fresh.p, rank.p, n, h->stream));
        h->launches += 2;
        ISL_LAUNCH(h, k_dg_write, h->grid_for(n, 256), 256, 0, first.p, rank.p, ne, nfaces, stride, L.total, L.begin[nf], next, d_elem_dof);
        uint32_t last_rank = 0, last_fresh = 0;
        ISL_CUDA(cudaMemcpyAsync(&last_rank, rank.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaMemcpyAsync(&last_fresh, fresh.p + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        next += ((int64_t)last_rank + last_fresh) * stride;
        if (getenv("ISL_VERBOSE"))
            fprintf(stderr, "[isl] dof generation on the device: n-face type %d, %lld items, %lld ids so far, %.2f ms since the start\n", nf, (long long)n,
                    (long long)next, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    }
    return next;
}
}  // namespace

int isl_dof_generate_device(isl_handle h, int fe_deg, int32_t* elem_dof, int64_t* n_obj) {
    return guarded([&] {
        ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(h->n_elems > 0, "mesh not set");
        ISL_REQUIRE(elem_dof != nullptr && n_obj != nullptr, "output arrays missing");
        flush_pending(h);
        const isl::FELayout L = isl::fe_layout(h->shape, fe_deg);
        const size_t n = (size_t)h->n_elems * L.total;
        cudaPointerAttributes at; std::memset(&at, 0, sizeof(at));
        const bool on_device = cudaPointerGetAttributes(&at, elem_dof) == cudaSuccess && at.type == cudaMemoryTypeDevice;
        cudaGetLastError();
        DevBuf<int32_t> buf;
        int32_t* d = elem_dof;
        if (!on_device) { buf.alloc(n); d = buf.p; }
        *n_obj = dof_generate_device(h, fe_deg, d);
        if (!on_device) {
            ISL_CUDA(cudaMemcpyAsync(elem_dof, d, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            ISL_CUDA(cudaStreamSynchronize(h->stream));
        }
    });
}

// ---- surface terms ----
namespace {
void assemble_neumann(isl_engine* h, int shape, int geom_deg, int64_t n_surf, const int32_t* domain_elem, const double* surf_x,
                      const double* surf_param, int quad_deg, int fe_deg, int ds, int t, const int32_t* rows, int mode, const double* data) {
    ISL_CUDA(cudaSetDevice(h->device));
    ISL_REQUIRE(h->n_eqn >= 0, "isl_system_create must be called first");
    ISL_REQUIRE(shape == ISL_TRI || shape == ISL_QUAD || shape == ISL_TET || shape == ISL_HEX, "surface terms need a 2-D or 3-D domain element");
    ISL_REQUIRE(mode == ISL_NEUMANN_CONSTANT || mode == ISL_NEUMANN_NORMAL || mode == ISL_NEUMANN_SAMPLED, "unknown surface force mode");
    ISL_REQUIRE(data != nullptr, "no surface force");
    ISL_REQUIRE(mode != ISL_NEUMANN_NORMAL || ds == isl::shape_dim(shape), "a force along the normal needs dof_size == dimension");
    require_live_system(h);
    flush_pending(h);
    if (n_surf <= 0) return;
    ISL_REQUIRE(surf_x != nullptr && surf_param != nullptr, "surface element arrays missing");
    const int dim = isl::shape_dim(shape), ld = dim - 1, sshape = isl::face_shape(shape);
    const isl::Basis sg(sshape, geom_deg), fe(shape, fe_deg);
    const isl::Rule R = isl::make_rule(sshape, quad_deg);
    const int P = sg.nfun, nq = R.n, nt = fe.nfun;
    std::vector<double> N((size_t)nq * P), dN((size_t)nq * P * ld);
    for (int q = 0; q < nq; q++) sg.eval(&R.p[(size_t)q * ld], &N[(size_t)q * P], &dN[(size_t)q * P * ld]);
    // distinct blocks of parameter coordinates and the test functions at xi(eta_q) for each of them
    // (SurfaceElement::localDomainCoordinate, base/mesh/SurfaceElement.hpp:42-55: xi = sum_p param_p N_p(eta))
    std::map<std::vector<double>, int> seen;
    std::vector<int32_t> pat((size_t)n_surf);
    std::vector<double> phi, block((size_t)P * dim), fun((size_t)nt);
    for (int64_t k = 0; k < n_surf; k++) {
        block.assign(surf_param + (size_t)k * P * dim, surf_param + (size_t)(k + 1) * P * dim);
        auto it = seen.find(block);
        if (it == seen.end()) {
            it = seen.emplace(block, (int)seen.size()).first;
            for (int q = 0; q < nq; q++) {
                double xi[3] = {0., 0., 0.};
                for (int p = 0; p < P; p++) for (int d = 0; d < dim; d++) xi[d] += block[(size_t)p * dim + d] * N[(size_t)q * P + p];
                fe.eval(xi, fun.data(), nullptr);
                phi.insert(phi.end(), fun.begin(), fun.end());
            }
        }
        pat[(size_t)k] = it->second;
    }
    NeumannParams p; std::memset(&p, 0, sizeof(p));
    DevBuf<double> d_sx, d_phi, d_dN, d_w, d_data; DevBuf<int32_t> d_pat, d_elem, d_rows;
    upload(h, d_sx, surf_x, (size_t)n_surf * P * dim); upload(h, d_phi, phi.data(), phi.size());
    upload(h, d_dN, dN.data(), dN.size()); upload(h, d_w, R.w.data(), (size_t)nq); upload(h, d_pat, pat.data(), pat.size());
    p.n_surf = n_surf; p.dim = dim; p.P = P; p.nq = nq; p.nt = nt; p.ds = ds; p.mode = mode;
    p.sx = d_sx.p; p.pat = d_pat.p; p.phi = d_phi.p; p.sdN = d_dN.p; p.w = d_w.p; p.rhs = h->rhs.p;
    if (mode == ISL_NEUMANN_SAMPLED) { upload(h, d_data, data, (size_t)n_surf * nq * ds); p.data = d_data.p; }
    else for (int c = 0; c < (mode == ISL_NEUMANN_NORMAL ? 1 : ds); c++) p.f[c] = data[c];
    if (rows) { upload(h, d_rows, rows, (size_t)n_surf * nt * ds); p.rows = d_rows.p; }
    else {
        FieldDev& f = h->fields[t];
        ISL_REQUIRE(domain_elem != nullptr, "domain elements of the surface elements missing");
        for (int64_t k = 0; k < n_surf; k++) ISL_REQUIRE(domain_elem[k] >= 0 && domain_elem[k] < h->n_elems, "surface element on an unknown domain element");
        upload(h, d_elem, domain_elem, (size_t)n_surf);
        p.elem = d_elem.p; p.ed = f.elem_dof.p; p.eqn = f.eqn.p;
        p.cptr = f.has_masters ? f.cptr.p : nullptr; p.cm = f.cmaster.p; p.cw = f.cweight.p;
    }
    const int64_t total = n_surf * nt * ds;
    if (dim == 3) ISL_LAUNCH(h, k_neumann<3>, h->grid_for(total, 128), 128, 0, p);
    else ISL_LAUNCH(h, k_neumann<2>, h->grid_for(total, 128), 128, 0, p);
    ISL_CUDA(cudaStreamSynchronize(h->stream));   // the staging buffers are released when this function returns
}
}  // namespace

int isl_boundary_surface(int shape, int geom_deg, int dim, const double* coords, const int32_t* conn, int64_t n_pairs,
                         const int64_t* pairs, int32_t* domain_elem, double* surf_x, double* surf_param, int* surf_shape,
                         int* nodes_per_surf) {
    return guarded([&] {
        ISL_REQUIRE(dim == isl::shape_dim(shape) && (dim == 2 || dim == 3), "surface elements of 2-D and 3-D meshes only");
        const int P = isl::boundary_surface(shape, geom_deg, dim, coords, conn, n_pairs, pairs, domain_elem, surf_x, surf_param);
        if (surf_shape) *surf_shape = isl::face_shape(shape);
        if (nodes_per_surf) *nodes_per_surf = P;
    });
}
int isl_surface_points(int surf_shape, int geom_deg, int dim, int64_t n_surf, const double* surf_x, int quad_deg, double* x,
                       double* normal, double* detg, int* nq) {
    return guarded([&] {
        ISL_REQUIRE(dim == isl::shape_dim(surf_shape) + 1, "surface shape and dimension do not match");
        const int n = isl::surface_points(surf_shape, geom_deg, dim, n_surf, surf_x, quad_deg, x, normal, detg);
        if (nq) *nq = n;
    });
}
int isl_assemble_neumann(isl_handle h, int64_t n_surf, const int32_t* domain_elem, const double* surf_x, const double* surf_param,
                         int quad_deg, int test_field, int mode, const double* data) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(test_field >= 0 && test_field < 5 && h->fields[test_field].set, "field not set");
        const FieldDev& f = h->fields[test_field];
        ISL_REQUIRE(h->n_elems > 0, "mesh not set");
        assemble_neumann(h, h->shape, h->geom_deg, n_surf, domain_elem, surf_x, surf_param, quad_deg, f.deg, f.ds, test_field, nullptr, mode, data);
    });
}
int isl_assemble_neumann_rows(isl_handle h, int shape, int geom_deg, int64_t n_surf, const double* surf_x, const double* surf_param,
                              int quad_deg, int fe_deg, int dof_size, const int32_t* rows, int mode, const double* data) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(rows != nullptr, "no equation numbers");
        ISL_REQUIRE(dof_size >= 1 && dof_size <= 3, "dof_size must be 1..3");
        assemble_neumann(h, shape, geom_deg, n_surf, nullptr, surf_x, surf_param, quad_deg, fe_deg, dof_size, -1, rows, mode, data);
    });
}

// host-side odd contributions (Neumann terms of the reference's own loops come through here element by element):
// no allocation and no wait per call; the operands are staged in a small arena that the stream reuses in order, a
// missing pattern entry is recorded on the device and reported by isl_finish
int isl_insert_lhs(isl_handle h, const double* mat, const int64_t* rows, int n_rows, const int64_t* cols, int n_cols) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        require_live_system(h);
        flush_pending(h);
        materialize_zero(h);
        h->val_is_zero = false;
        ISL_REQUIRE(h->nnz > 0, "no pattern registered");
        for (int i = 0; i < n_rows; i++) ISL_REQUIRE(rows[i] >= 0 && rows[i] < h->n_eqn, "Row index out of bound: " + std::to_string(rows[i]));
        for (int j = 0; j < n_cols; j++) ISL_REQUIRE(cols[j] >= 0 && cols[j] < h->n_eqn, "Col index out of bound: " + std::to_string(cols[j]));
        const size_t nm = (size_t)n_rows * n_cols;
        char* a = h->insert_arena(nm * 8 + (size_t)(n_rows + n_cols) * 8);
        double* dm = reinterpret_cast<double*>(a);
        int64_t* dr = reinterpret_cast<int64_t*>(a + nm * 8);
        int64_t* dc = dr + n_rows;
        ISL_CUDA(cudaMemcpyAsync(dm, mat, nm * 8, cudaMemcpyDefault, h->stream));
        ISL_CUDA(cudaMemcpyAsync(dr, rows, (size_t)n_rows * 8, cudaMemcpyDefault, h->stream));
        ISL_CUDA(cudaMemcpyAsync(dc, cols, (size_t)n_cols * 8, cudaMemcpyDefault, h->stream));
        const int n = (int)nm;
        ISL_LAUNCH(h, k_insert_lhs, (n + 127) / 128, 128, 0, dm, dr, n_rows, dc, n_cols, h->rowptr.p, h->col.p, h->val.p, h->ins_err.p);
    });
}
int isl_insert_rhs(isl_handle h, const double* vec, const int64_t* rows, int n_rows) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        require_live_system(h);
        flush_pending(h);
        for (int i = 0; i < n_rows; i++) ISL_REQUIRE(rows[i] >= 0 && rows[i] < h->n_eqn, std::to_string(rows[i]) + " out of bound");
        char* a = h->insert_arena((size_t)n_rows * 16);
        double* dv = reinterpret_cast<double*>(a);
        int64_t* dr = reinterpret_cast<int64_t*>(a + (size_t)n_rows * 8);
        ISL_CUDA(cudaMemcpyAsync(dv, vec, (size_t)n_rows * 8, cudaMemcpyDefault, h->stream));
        ISL_CUDA(cudaMemcpyAsync(dr, rows, (size_t)n_rows * 8, cudaMemcpyDefault, h->stream));
        ISL_LAUNCH(h, k_insert_rhs, (n_rows + 127) / 128, 128, 0, dv, dr, n_rows, h->rhs.p);
    });
}

int isl_finish(isl_handle h, int64_t* n_eqn, int64_t* nnz) {
    return guarded([&] {
        ISL_REQUIRE(!h->sys_stale, "mesh or field arrays were replaced after assembly into this system had started: create a new solver first");
        flush_pending(h);
        materialize_zero(h);
        ISL_CUDA(cudaSetDevice(h->device));
        if (h->sys_pairs != h->pattern_pairs && !h->sys_pairs.empty()) {
            // the cached pattern holds blocks this system never registered: rebuild exactly
            build_pattern(h, h->sys_pairs);
        }
        if (h->ins_err.p) {
            int err = 0;
            ISL_CUDA(cudaMemcpyAsync(&err, h->ins_err.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            ISL_CUDA(cudaStreamSynchronize(h->stream));
            ISL_REQUIRE(!err, "TripletContainer had not been properly set up");
        }
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        if (n_eqn) *n_eqn = h->n_eqn;
        if (nnz) *nnz = h->sys_pairs.empty() ? 0 : h->nnz;
    });
}
int isl_get_csr(isl_handle h, int64_t* rowptr, int32_t* col, double* val, double* rhs) {
    return guarded([&] {
        flush_pending(h);
        materialize_zero(h);
        ISL_CUDA(cudaSetDevice(h->device));
        if (rowptr) {
            if (h->rowptr.p) ISL_CUDA(cudaMemcpyAsync(rowptr, h->rowptr.p, (h->n_eqn + 1) * sizeof(int64_t), cudaMemcpyDefault, h->stream));
            else ISL_CUDA(cudaMemsetAsync(rowptr, 0, (h->n_eqn + 1) * sizeof(int64_t), h->stream));
        }
        if (col && h->nnz) ISL_CUDA(cudaMemcpyAsync(col, h->col.p, h->nnz * sizeof(int32_t), cudaMemcpyDefault, h->stream));
        if (val && h->nnz) ISL_CUDA(cudaMemcpyAsync(val, h->val.p, h->nnz * sizeof(double), cudaMemcpyDefault, h->stream));
        if (rhs && h->n_eqn) ISL_CUDA(cudaMemcpyAsync(rhs, h->rhs.p, h->n_eqn * sizeof(double), cudaMemcpyDefault, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
    });
}
int isl_get_csr_async(isl_handle h, double* val, double* rhs) {
    return guarded([&] {
        ISL_REQUIRE(!h->sys_stale && !h->sys_released, "no finished system to hand over");
        ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);
        materialize_zero(h);
        if (!h->copy_stream) {
            ISL_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
            ISL_CUDA(cudaEventCreateWithFlags(&h->ev_asm, cudaEventDisableTiming));
            ISL_CUDA(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
        }
        if (h->val_alt.n != (size_t)h->nnz) h->val_alt.alloc((size_t)h->nnz);
        if (h->rhs_alt.n != (size_t)h->n_eqn) h->rhs_alt.alloc((size_t)h->n_eqn);
        // the buffers that come back into use were read by the previous asynchronous copy: the engine stream waits for it
        if (h->copy_in_flight) ISL_CUDA(cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
        ISL_CUDA(cudaEventRecord(h->ev_asm, h->stream));
        h->val.swap(h->val_alt); h->rhs.swap(h->rhs_alt);
        ISL_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_asm, 0));
        if (val && h->nnz) ISL_CUDA(cudaMemcpyAsync(val, h->val_alt.p, h->nnz * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
        if (rhs && h->n_eqn) ISL_CUDA(cudaMemcpyAsync(rhs, h->rhs_alt.p, h->n_eqn * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
        ISL_CUDA(cudaEventRecord(h->ev_copy, h->copy_stream));
        h->copy_in_flight = true;
        h->sys_released = true;   // the system now lives in the copy buffers: isl_system_create starts the next one
    });
}
int isl_copy_wait(isl_handle h) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        if (h->copy_in_flight) ISL_CUDA(cudaEventSynchronize(h->ev_copy));
    });
}
int isl_get_device_csr(isl_handle h, int64_t** rowptr, int32_t** col, double** val, double** rhs) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);
        materialize_zero(h);
        if (rowptr) *rowptr = h->rowptr.p;
        if (col) *col = h->col.p;
        if (val) *val = h->val.p;
        if (rhs) *rhs = h->rhs.p;
    });
}
int isl_rhs_value(isl_handle h, int64_t index, double* value) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);
        ISL_REQUIRE(index >= 0 && index < h->n_eqn, "index out of bound");
        ISL_CUDA(cudaMemcpyAsync(value, h->rhs.p + index, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
    });
}
int isl_rhs_norm(isl_handle h, double* norm) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);
        ISL_REQUIRE(h->n_eqn > 0, "empty system");
        h->scratch_d.alloc(1);
        ISL_CUDA(cudaMemsetAsync(h->scratch_d.p, 0, sizeof(double), h->stream));
        ISL_LAUNCH(h, k_sumsq, h->grid_for(h->n_eqn, 256), 256, 0, h->rhs.p, h->n_eqn, h->scratch_d.p);
        double s = 0.;
        ISL_CUDA(cudaMemcpyAsync(&s, h->scratch_d.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        *norm = std::sqrt(s) / (double)h->n_eqn;  // Eigen3.hpp:133-138 divides by the segment length
    });
}

int isl_solve_cg(isl_handle h, double tol, int64_t max_iter, int64_t* iterations, double* error) {
    return guarded([&] {
        flush_pending(h);
        materialize_zero(h);
        ISL_CUDA(cudaSetDevice(h->device));
        const int64_t it = solve_cg(h, tol, max_iter, error);
        if (iterations) *iterations = it;
    });
}

/* base::dof::setDoFsFromSolver / addToDoFsFromSolver (base/dof/Distribute.hpp:35-56,139-215) on the device */
int isl_distribute(isl_handle h, int field, int add) {
    return guarded([&] {
        ISL_REQUIRE(field >= 0 && field < 5 && h->fields[field].set, "field not set");
        ISL_REQUIRE(h->n_eqn >= 0 && h->rhs.p, "no system");
        ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);
        FieldDev& f = h->fields[field];
        const int64_t n = f.n_obj * f.ds;
        if (n == 0) return;
        ISL_LAUNCH(h, k_distribute_active, h->grid_for(n, 256), 256, 0, f.eqn.p, h->rhs.p, f.values.p, n, add ? 1 : 0);
        DevBuf<int32_t> map; DevBuf<int> derr;
        derr.alloc(1);
        ISL_CUDA(cudaMemsetAsync(derr.p, 0, sizeof(int), h->stream));
        if (f.has_masters) {
            map.alloc((size_t)std::max<int64_t>(h->n_eqn, 1));
            ISL_CUDA(cudaMemsetAsync(map.p, 0xff, map.n * sizeof(int32_t), h->stream));
            ISL_LAUNCH(h, k_eqn2dof, h->grid_for(n, 256), 256, 0, f.eqn.p, n, map.p);
        }
        ISL_LAUNCH(h, k_distribute_constrained, h->grid_for(n, 256), 256, 0, f.status.p, f.presc.p, f.has_masters ? f.cptr.p : nullptr,
                   f.cmaster.p, f.cweight.p, map.p, 0, (int32_t)h->n_eqn, f.values.p, n, derr.p);
        int err = 0;
        ISL_CUDA(cudaMemcpyAsync(&err, derr.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        ISL_REQUIRE(!err, "a master DoF of a linear constraint belongs to another field (not supported by isl_distribute)");
    });
}
int isl_field_get_values(isl_handle h, int field, double* values) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(field >= 0 && field < 5 && h->fields[field].set, "field not set");
        ISL_REQUIRE(values, "null output");
        flush_pending(h);
        FieldDev& f = h->fields[field];
        ISL_CUDA(cudaMemcpyAsync(values, f.values.p, (size_t)f.n_obj * f.ds * sizeof(double), cudaMemcpyDefault, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int isl_measure_fp64_peak(isl_handle h, double* tflops) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(tflops, "null output");
        flush_pending(h);
        DevBuf<double> out; out.alloc(1);
        cudaEvent_t a, b;
        ISL_CUDA(cudaEventCreate(&a)); ISL_CUDA(cudaEventCreate(&b));
        constexpr int CH = 8;
        const int grid = h->n_sm * 8, iters = 2048;
        double best = 0.;
        for (int rep = 0; rep < 4; rep++) {
            ISL_CUDA(cudaEventRecord(a, h->stream));
            ISL_LAUNCH(h, k_dfma_peak<CH>, grid, 256, 0, out.p, iters, 0.999999, 1e-9);
            ISL_CUDA(cudaEventRecord(b, h->stream));
            ISL_CUDA(cudaEventSynchronize(b));
            float ms = 0.f;
            ISL_CUDA(cudaEventElapsedTime(&ms, a, b));
            const double fl = 2.0 * CH * 16.0 * iters * 256.0 * grid;
            if (rep > 0 && ms > 0.f) best = std::max(best, fl / (ms * 1e-3) / 1e12);
        }
        cudaEventDestroy(a); cudaEventDestroy(b);
        *tflops = best;
    });
}

int isl_measure_red_peak(isl_handle h, int pattern, double* gatomics_per_s) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        ISL_REQUIRE(gatomics_per_s && pattern >= 0 && pattern <= 2, "bad arguments");
        flush_pending(h);
        const unsigned long long n = 1ull << 28;   // 2 GiB of doubles: far larger than L2
        DevBuf<double> a; a.alloc((size_t)n);
        ISL_CUDA(cudaMemsetAsync(a.p, 0, (size_t)n * sizeof(double), h->stream));
        cudaEvent_t e0, e1;
        ISL_CUDA(cudaEventCreate(&e0)); ISL_CUDA(cudaEventCreate(&e1));
        const int grid = h->n_sm * 8, iters = 512;
        double best = 0.;
        for (int rep = 0; rep < 3; rep++) {
            ISL_CUDA(cudaEventRecord(e0, h->stream));
            ISL_LAUNCH(h, k_red_peak, grid, 256, 0, a.p, n, pattern, iters);
            ISL_CUDA(cudaEventRecord(e1, h->stream));
            ISL_CUDA(cudaEventSynchronize(e1));
            float ms = 0.f;
            ISL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            const double ops = (double)grid * 256.0 * iters * (pattern == 1 ? 3.0 : 1.0);
            if (rep > 0 && ms > 0.f) best = std::max(best, ops / (ms * 1e-3) / 1e9);
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        *gatomics_per_s = best;
    });
}

// ---- multi-GPU: communicator and interface-row exchange inside the engine (isl_comm.cuh) ----
int isl_comm_unique_id(void* id128) {
    return guarded([&] {
        ISL_REQUIRE(id128, "null id buffer");
        nccl_dyn::ncclUniqueId id;
        ISL_NCCL(nccl_dyn::api().GetUniqueId(&id));
        std::memcpy(id128, &id, sizeof(id));
    });
}
int isl_comm_init(isl_handle h, const void* id128, int rank, int world) {
    return guarded([&] {
        ISL_REQUIRE(id128 && world >= 1 && rank >= 0 && rank < world, "bad communicator arguments");
        ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);
        auto c = std::make_unique<CommState>();
        c->rank = rank; c->world = world;
        nccl_dyn::ncclUniqueId id;
        std::memcpy(&id, id128, sizeof(id));
        ISL_NCCL(nccl_dyn::api().CommInitRank(&c->comm, world, id, rank));
        int prio_lo = 0, prio_hi = 0;
        ISL_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        ISL_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));   // its blocks are dispatched first
        ISL_CUDA(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
        ISL_CUDA(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
        ISL_CUDA(cudaEventCreateWithFlags(&c->ev_iface, cudaEventDisableTiming));
        h->comm = std::move(c);
    });
}
int isl_comm_destroy(isl_handle h) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        if (!h->comm) return;
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->comm->stream));
        h->comm.reset();
    });
}
int isl_exchange_setup(isl_handle h, int64_t n_local, const int64_t* l2g, int64_t own_lo, int64_t own_hi, int n_seg,
                       const int* seg_owner, const int64_t* seg_lo, const int64_t* seg_hi) {
    return guarded([&] {
        ISL_REQUIRE(h->comm, "isl_comm_init must be called first");
        ISL_REQUIRE(n_local == h->n_eqn, "l2g must cover the local system");
        ISL_REQUIRE(h->nnz > 0 && h->rowptr.p, "register the pattern before the exchange plan");
        ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);
        CommState& c = *h->comm;
        auto& N = nccl_dyn::api();
        const int W = c.world;
        c.plan = false; c.sends.clear(); c.recvs.clear(); c.iface_event_valid = false;
        upload(h, h->l2g, l2g, (size_t)n_local);
        c.iface_row.alloc((size_t)std::max<int64_t>(n_local, 1));
        ISL_CUDA(cudaMemsetAsync(c.iface_row.p, 0, (size_t)std::max<int64_t>(n_local, 1), h->stream));
        // what I send to whom: (entries, rows) per destination; everybody learns the whole table
        std::vector<int64_t> mine((size_t)W * 2, 0);
        for (int k = 0; k < n_seg; k++) {
            ISL_REQUIRE(seg_owner[k] >= 0 && seg_owner[k] < W && seg_owner[k] != c.rank, "bad ghost segment owner");
            ISL_REQUIRE(seg_lo[k] >= 0 && seg_lo[k] <= seg_hi[k] && seg_hi[k] <= n_local, "bad ghost segment");
            ISL_REQUIRE(mine[(size_t)seg_owner[k] * 2 + 1] == 0, "one ghost segment per owner");
            int64_t rp[2];
            ISL_CUDA(cudaMemcpyAsync(&rp[0], h->rowptr.p + seg_lo[k], sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
            ISL_CUDA(cudaMemcpyAsync(&rp[1], h->rowptr.p + seg_hi[k], sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
            ISL_CUDA(cudaStreamSynchronize(h->stream));
            if (seg_hi[k] == seg_lo[k]) continue;
            c.sends.push_back(CommSend{seg_owner[k], rp[0], rp[1] - rp[0], seg_lo[k], seg_hi[k] - seg_lo[k]});
            mine[(size_t)seg_owner[k] * 2] = rp[1] - rp[0]; mine[(size_t)seg_owner[k] * 2 + 1] = seg_hi[k] - seg_lo[k];
            ISL_LAUNCH(h, k_comm_mark, h->grid_for(seg_hi[k] - seg_lo[k], 256), 256, 0, c.iface_row.p, seg_lo[k], seg_hi[k]);
        }
        DevBuf<int64_t> dmine, dtable;
        upload_vec(h, dmine, mine); dtable.alloc((size_t)W * W * 2);
        ISL_NCCL(N.AllGather(dmine.p, dtable.p, (size_t)W * 2, nccl_dyn::ncclInt64, c.comm, h->stream));
        std::vector<int64_t> table((size_t)W * W * 2);
        ISL_CUDA(cudaMemcpyAsync(table.data(), dtable.p, table.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        // keys of my ghost entries, global ids of my ghost rows
        std::vector<std::unique_ptr<DevBuf<uint64_t>>> skeys;
        for (const CommSend& sd : c.sends) {
            auto kb = std::make_unique<DevBuf<uint64_t>>(); kb->alloc((size_t)std::max<int64_t>(sd.n_val, 1));
            ISL_LAUNCH(h, k_comm_keys, (unsigned)std::min<int64_t>(sd.n_rows, 65535), 32, 0, h->rowptr.p, h->col.p, h->l2g.p, sd.row_lo,
                       sd.row_lo + sd.n_rows, kb->p);
            skeys.push_back(std::move(kb));
        }
        std::vector<std::unique_ptr<DevBuf<uint64_t>>> rkeys;
        std::vector<std::unique_ptr<DevBuf<int64_t>>> rrows;
        for (int src = 0; src < W; src++) {
            const int64_t ne = table[((size_t)src * W + c.rank) * 2], nr = table[((size_t)src * W + c.rank) * 2 + 1];
            if (src == c.rank || nr == 0) continue;
            auto rv = std::make_unique<CommRecv>();
            rv->src = src; rv->n_val = ne; rv->n_rows = nr;
            rv->pos.alloc((size_t)std::max<int64_t>(ne, 1)); rv->rows.alloc((size_t)nr);
            rv->bval.alloc((size_t)std::max<int64_t>(ne, 1)); rv->brhs.alloc((size_t)nr);
            auto kb = std::make_unique<DevBuf<uint64_t>>(); kb->alloc((size_t)std::max<int64_t>(ne, 1));
            auto rb = std::make_unique<DevBuf<int64_t>>(); rb->alloc((size_t)nr);
            rkeys.push_back(std::move(kb)); rrows.push_back(std::move(rb));
            c.recvs.push_back(std::move(rv));
        }
        ISL_NCCL(N.GroupStart());
        for (size_t k = 0; k < c.sends.size(); k++) {
            const CommSend& sd = c.sends[k];
            ISL_NCCL(N.Send(skeys[k]->p, (size_t)sd.n_val, nccl_dyn::ncclUint64, sd.dst, c.comm, h->stream));
            ISL_NCCL(N.Send(h->l2g.p + sd.row_lo, (size_t)sd.n_rows, nccl_dyn::ncclInt64, sd.dst, c.comm, h->stream));
        }
        for (size_t k = 0; k < c.recvs.size(); k++) {
            ISL_NCCL(N.Recv(rkeys[k]->p, (size_t)c.recvs[k]->n_val, nccl_dyn::ncclUint64, c.recvs[k]->src, c.comm, h->stream));
            ISL_NCCL(N.Recv(rrows[k]->p, (size_t)c.recvs[k]->n_rows, nccl_dyn::ncclInt64, c.recvs[k]->src, c.comm, h->stream));
        }
        ISL_NCCL(N.GroupEnd());
        DevBuf<int> derr; derr.alloc(1);
        ISL_CUDA(cudaMemsetAsync(derr.p, 0, sizeof(int), h->stream));
        for (size_t k = 0; k < c.recvs.size(); k++) {
            CommRecv& rv = *c.recvs[k];
            if (rv.n_val) ISL_LAUNCH(h, k_comm_locate, h->grid_for(rv.n_val, 128), 128, 0, rkeys[k]->p, rv.n_val, h->l2g.p, own_lo, own_hi,
                                     h->rowptr.p, h->col.p, rv.pos.p, derr.p);
            ISL_LAUNCH(h, k_comm_locate_rows, h->grid_for(rv.n_rows, 128), 128, 0, rrows[k]->p, rv.n_rows, h->l2g.p, own_lo, own_hi, rv.rows.p,
                       c.iface_row.p, derr.p);
        }
        int err = 0;
        ISL_CUDA(cudaMemcpyAsync(&err, derr.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        ISL_CUDA(cudaStreamSynchronize(h->stream));
        ISL_REQUIRE(!err, "ghost entry missing in the owner's pattern (halo elements not registered?) or row not owned");
        c.plan = true; c.plan_nnz = h->nnz;
        for (auto& kv : h->patchsets) { kv.second->perm_built = false; kv.second->perm.release(); kv.second->n_iface = 0; }
    });
}
int isl_exchange(isl_handle h) {
    return guarded([&] {
        ISL_REQUIRE(h->comm && h->comm->plan, "isl_exchange_setup must be called first");
        ISL_REQUIRE(h->comm->plan_nnz == h->nnz, "the pattern changed after isl_exchange_setup");
        ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);     // may be the split launch that records ev_iface
        materialize_zero(h);
        CommState& c = *h->comm;
        auto& N = nccl_dyn::api();
        if (c.iface_event_valid) {
            // the interface patches were launched on the communication stream itself: stream order is enough
        } else {
            ISL_CUDA(cudaEventRecord(c.ev_ready, h->stream));
            ISL_CUDA(cudaStreamWaitEvent(c.stream, c.ev_ready, 0));
        }
        c.iface_event_valid = false;
        ISL_NCCL(N.GroupStart());
        for (const CommSend& sd : c.sends) {
            if (sd.n_val) ISL_NCCL(N.Send(h->val.p + sd.val_lo, (size_t)sd.n_val, nccl_dyn::ncclFloat64, sd.dst, c.comm, c.stream));
            ISL_NCCL(N.Send(h->rhs.p + sd.row_lo, (size_t)sd.n_rows, nccl_dyn::ncclFloat64, sd.dst, c.comm, c.stream));
        }
        for (auto& rv : c.recvs) {
            if (rv->n_val) ISL_NCCL(N.Recv(rv->bval.p, (size_t)rv->n_val, nccl_dyn::ncclFloat64, rv->src, c.comm, c.stream));
            ISL_NCCL(N.Recv(rv->brhs.p, (size_t)rv->n_rows, nccl_dyn::ncclFloat64, rv->src, c.comm, c.stream));
        }
        ISL_NCCL(N.GroupEnd());
        for (auto& rv : c.recvs) {
            if (rv->n_val) {
                k_unpack_add<<<h->grid_for(rv->n_val, 256), 256, 0, c.stream>>>(h->val.p, rv->pos.p, rv->n_val, rv->bval.p);
                h->launches++;
            }
            k_unpack_add<<<h->grid_for(rv->n_rows, 256), 256, 0, c.stream>>>(h->rhs.p, rv->rows.p, rv->n_rows, rv->brhs.p);
            h->launches++;
            ISL_CUDA(cudaGetLastError());
        }
        c.join_pending = true;
        comm_join(h);
        h->val_is_zero = false;
    });
}

int isl_pack_entries(isl_handle h, int which, const int64_t* idx_dev, int64_t n, double* out_dev) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);
        materialize_zero(h);
        if (n == 0) return;
        const double* src = which == 1 ? h->rhs.p : h->val.p;
        ISL_LAUNCH(h, k_pack, h->grid_for(n, 256), 256, 0, src, idx_dev, n, out_dev);
    });
}
int isl_unpack_add_entries(isl_handle h, int which, const int64_t* idx_dev, int64_t n, const double* in_dev) {
    return guarded([&] { ISL_CUDA(cudaSetDevice(h->device));
        flush_pending(h);
        materialize_zero(h);
        if (n == 0) return;
        double* dst = which == 1 ? h->rhs.p : h->val.p;
        ISL_LAUNCH(h, k_unpack_add, h->grid_for(n, 256), 256, 0, dst, idx_dev, n, in_dev);
    });
}

}  // extern "C"
