// =============================================================================
// isl_gather.cuh -- atomic-free assembly of the hyperelastic tangent: element matrices to memory, CSR rows gathered
// (included by isl_engine.cu).
//
// Measured (profiles/r2/session30.log): FP64 atomic adds without return reach 308 G/s when consecutive threads add to
// consecutive entries, but only 39.5 G/s as triples at scattered places -- which is what a local matrix scattered through a
// slot map produces (6 561 adds per Q2 x 3 element into 2 187 different 32-byte sectors); k_tangent_hypel_sym with the
// atomic scatter sits at 0.88 of that rate.  So the scatter is turned around, as for the Q1 row kernels:
//   A. k_tangent_hypel_sym writes the finished local matrix K_e (it already assembles it in shared memory) to
//      Kbuf[e][nr][nr] with coalesced stores;
//   B. k_gather_rows: one WARP per CSR row r.  The (element, local row) pairs that contribute to r are listed once
//      (stable radix sort of the equation numbers of all local rows: the visiting order of the reference's element loop is
//      kept, so the sum is deterministic); for every pair the warp reads the local row (nr contiguous doubles) and adds
//      it into a row buffer in shared memory at the positions pos[pair][j] (uint16, relative to the row start; the columns
//      of one element are distinct, so the lanes never collide), Dirichlet columns go into the lift of rhs[r]; the row
//      leaves with coalesced plain stores (or read-add-store when the system already holds other contributions).
// Traffic per element (Q2 x 3): 52 KB K_e written + read, 13 KB positions, its share of the CSR values: no atomics, no
// slot map.  Reference semantics unchanged: asmb/assembleMatrix.hpp:56-130 (ACTIVE x ACTIVE -> matrix, ACTIVE x
// CONSTRAINED -> rhs -= g K).
// =============================================================================
#pragma once

struct GatherSet {
    bool ok = false;
    int nr = 0; int64_t n_pairs = 0, n_rows = 0; int max_len = 0;
    DevBuf<int32_t> pair;        // sorted position -> element * nr + local row
    DevBuf<int64_t> row_start;   // [n_rows + 1] into pair
    DevBuf<uint16_t> pos;        // [n_pairs][nr] position of local column j inside the CSR row, 0xffff: column not ACTIVE
    DevBuf<double> Kbuf;         // [n_elems][nr][nr]
};

struct GatherParams {
    const int32_t* pair; const int64_t* row_start; const uint16_t* pos; const double* Kbuf;
    int nr, nt, ds; int64_t n_rows;
    const int64_t* rowptr; double* val; double* rhs;
    const int32_t* ed; const uint8_t* status; const double* presc; const double* values; int incremental;
    int store;                   // 1: the rows hold nothing yet and nothing else contributes: plain stores
    int buf_len;                 // row buffer per warp (doubles)
};

__global__ void k_gs_keys(const int32_t* elem_eqn, int64_t n, int32_t* key, int32_t* idx) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int32_t q = elem_eqn[p];
        key[p] = q >= 0 ? q : 0x7fffffff;
        idx[p] = (int32_t)p;
    }
}
// first sorted position whose key is >= r, for r = 0 .. n_rows
__global__ void k_gs_row_start(const int32_t* sorted_key, int64_t n, int64_t n_rows, int64_t* row_start) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = n;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)sorted_key[mid] < r) lo = mid + 1; else hi = mid; }
        row_start[r] = lo;
    }
}
__global__ void k_gs_pos(const int32_t* pair, const int32_t* sorted_key, int64_t n_pairs, int nr, const int32_t* elem_eqn,
                         const int64_t* rowptr, const int32_t* col, uint16_t* pos, int* err) {
    const int64_t n = n_pairs * nr;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = t / nr; const int j = (int)(t - q * nr);
        const int64_t e = pair[q] / nr;
        const int32_t r = sorted_key[q], c = elem_eqn[e * nr + j];
        uint16_t v = 0xffff;
        if (c >= 0) {
            const int64_t at = isl_find_in_row(rowptr, col, r, c);
            if (at < 0 || at - rowptr[r] >= 0xfffe) err[0] = 1; else v = (uint16_t)(at - rowptr[r]);
        }
        pos[t] = v;
    }
}
// two local columns of one element with the same equation number (tied numbering) would collide in the row buffer
__global__ void k_gs_dup(const int32_t* elem_eqn, int64_t n_elems, int nr, int* err) {
    const int64_t n = n_elems * nr;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / nr; const int j = (int)(t - e * nr);
        const int32_t c = elem_eqn[t];
        if (c < 0) continue;
        for (int j2 = 0; j2 < j; j2++) if (elem_eqn[e * nr + j2] == c) err[0] = 1;
    }
}
__global__ void k_gs_max_len(const int64_t* rowptr, int64_t n_rows, int* max_len) {
    int m = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) m = max(m, (int)(rowptr[r + 1] - rowptr[r]));
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_len, m);
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_gather_rows(const GatherParams p) {
    extern __shared__ double gsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* buf = gsm + (size_t)warp * p.buf_len;
    for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < p.n_rows; r += (int64_t)gridDim.x * WARPS) {
        const int64_t s = p.rowptr[r];
        const int len = (int)(p.rowptr[r + 1] - s);
        const int64_t q0 = p.row_start[r], q1 = p.row_start[r + 1];
        if (q0 == q1 && !p.store) continue;
        for (int k = lane; k < len; k += 32) buf[k] = 0.;
        double lift = 0.;
        __syncwarp();
        if (p.nr <= 96) {
            // software pipeline over the (element, local row) pairs: the loads of pair q + 1 are in flight while pair q is
            // added into the row buffer (up to three entries per lane)
            uint16_t at_n[3]; double v_n[3]; int64_t e_n = 0;
            auto fetch = [&](int64_t q) {
                const int32_t pr = __ldg(p.pair + q);
                e_n = pr / p.nr;
                const int i = pr - (int)e_n * p.nr;
                const double* krow = p.Kbuf + ((size_t)e_n * p.nr + i) * p.nr;
                const uint16_t* pq = p.pos + (size_t)q * p.nr;
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    const int j = lane + 32 * u;
                    const bool in = j < p.nr;
                    at_n[u] = in ? __ldg(pq + j) : (uint16_t)0xfffe;
                    v_n[u] = in ? __ldg(krow + j) : 0.;
                }
            };
            if (q0 < q1) fetch(q0);
            for (int64_t q = q0; q < q1; q++) {
                uint16_t at[3]; double v[3]; const int64_t e = e_n;
#pragma unroll
                for (int u = 0; u < 3; u++) { at[u] = at_n[u]; v[u] = v_n[u]; }
                if (q + 1 < q1) fetch(q + 1);
#pragma unroll
                for (int u = 0; u < 3; u++) {
                    if (at[u] < 0xfffe) buf[at[u]] += v[u];
                    else if (at[u] == 0xffff) {   // column not ACTIVE: the Dirichlet lift of a CONSTRAINED DoF, nothing for an inactive one
                        const int j = lane + 32 * u;
                        const size_t k = (size_t)p.ed[e * p.nt + j / p.ds] * p.ds + (j % p.ds);
                        if (p.status[k] == ISL_CONSTRAINED) lift += (p.incremental ? p.presc[k] - p.values[k] : p.presc[k]) * v[u];
                    }
                }
                __syncwarp();
            }
        } else
        for (int64_t q = q0; q < q1; q++) {
            const int32_t pr = __ldg(p.pair + q);
            const int64_t e = pr / p.nr; const int i = pr - (int)e * p.nr;
            const double* krow = p.Kbuf + ((size_t)e * p.nr + i) * p.nr;
            const uint16_t* pq = p.pos + (size_t)q * p.nr;
            for (int j = lane; j < p.nr; j += 32) {
                const uint16_t at = __ldg(pq + j);
                const double v = __ldg(krow + j);
                if (at != 0xffff) buf[at] += v;
                else {   // column not ACTIVE: the Dirichlet lift of a CONSTRAINED DoF, nothing for an inactive one
                    const size_t k = (size_t)p.ed[e * p.nt + j / p.ds] * p.ds + (j % p.ds);
                    if (p.status[k] == ISL_CONSTRAINED) lift += (p.incremental ? p.presc[k] - p.values[k] : p.presc[k]) * v;
                }
            }
            __syncwarp();
        }
        if (p.store) { for (int k = lane; k < len; k += 32) p.val[s + k] = buf[k]; }
        else { for (int k = lane; k < len; k += 32) p.val[s + k] += buf[k]; }
        for (int o = 16; o > 0; o >>= 1) lift += __shfl_xor_sync(0xffffffffu, lift, o);
        if (lane == 0 && lift != 0.) atomicAdd(p.rhs + r, -lift);
        __syncwarp();
    }
}

// =============================================================================
// The same turn-around for the GENERIC kernels (k_tangent: Laplace with constant / sampled conductivity, vector Laplace,
// Mass, Convection, PressureGradient, VelocityDivergence; config 5 = three of them on Taylor-Hood tetrahedra), where test
// and trial field differ and the local matrix is small:
//   * k_tangent stores its local matrix instead of scattering it: Kgen[e][KR][KC], with KR x KC = nt x nc for the
//     integrands that repeat one scalar entry on every DoF component (entry (M,c; N,c) = K[M][N]: "compact", 100 doubles
//     instead of 900 for the P2 velocity block) and (nt dst) x (nc dsc) for the two Stokes coupling blocks;
//   * k_gen_gather_rows: a SUB-WARP of G = 8, 16 or 32 lanes per CSR row (KC is 4 .. 30 here; a whole warp per row would
//     leave most lanes idle), the (element, local row) pairs of the TEST field listed once per field (RowPairs), the
//     positions of the KC local columns of the TRIAL field inside the row per pair (uint16).  Sum order = element order
//     of the caller, so the result is reproducible bit for bit (the atomic scatter is not).
// Reference semantics: asmb/assembleMatrix.hpp:56-130, fluid/PressureGradient.hpp:76-112, VelocityDivergence.hpp:67-124.
// =============================================================================
struct RowPairs {                // per test field
    bool ok = false;
    int nr = 0; int64_t n_pairs = 0, n_rows = 0;
    DevBuf<int32_t> pair;        // sorted position -> element * nr + local row
    DevBuf<int64_t> row_start;   // [n_rows + 1] into pair
};
struct GenGatherSet {            // per (test field, trial field, compact)
    bool ok = false;
    int KR = 0, KC = 0, compact = 0, max_len = 0;
    DevBuf<uint16_t> pos;        // [n_pairs][KC]
    DevBuf<int32_t> krow;        // [n_pairs] row of Kgen the pair reads: element * KR + local row (or local node, compact)
    // the columns a row receives from THIS block lie in [lo, hi] of the row (the trial field's equations; a few pressure
    // columns at the end of a velocity row for the Stokes coupling block): the row buffer covers that range only
    DevBuf<int32_t> row_lo, row_hi; int max_width = 0;
};
struct GenGatherParams {
    const int32_t* pair; const int64_t* row_start; const uint16_t* pos; const double* Kbuf;
    int nr, dst, KR, KC, compact, nc, dsc; int64_t n_rows;
    const int64_t* rowptr; double* val; double* rhs;
    const int32_t* ed_c; const uint8_t* st_c; const double* presc_c; const double* val_c; int incremental;
    int store;                   // 1: the system holds nothing yet: every row is written completely (no memset needed)
    int buf_len;                 // row buffer per sub-warp (doubles)
    const int32_t* row_lo; const int32_t* row_hi;   // nullptr: the buffer covers the whole row
    const int32_t* krow;         // nullptr: derived from pair[] with two integer divisions per pair
};

__global__ void k_gg_range_init(int32_t* lo, int32_t* hi, int64_t n) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) { lo[r] = 0x7fffffff; hi[r] = -1; }
}
__global__ void k_gg_max_width(const int32_t* lo, const int32_t* hi, int64_t n, int* out) {
    int m = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) m = max(m, hi[r] - lo[r] + 1);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__global__ void k_gg_pos(const int32_t* pair, int64_t n_pairs, int nr, int dst, int KC, int compact, int dsc, int ncl,
                         const int32_t* elem_eqn_t, const int32_t* elem_eqn_c, const int64_t* rowptr, const int32_t* col,
                         uint16_t* pos, int32_t* row_lo, int32_t* row_hi, int32_t* krow, int KR, int* err) {
    const int64_t n = n_pairs * KC;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = t / KC; const int j = (int)(t - q * KC);
        const int32_t pr = pair[q];
        const int64_t e = pr / nr; const int i = pr - (int)e * nr;
        if (j == 0) krow[q] = (int32_t)(e * KR + (compact ? i / dst : i));
        const int jc = compact ? j * dsc + (i % dst) : j;
        const int32_t r = elem_eqn_t[pr], c = elem_eqn_c[e * ncl + jc];
        uint16_t v = 0xffff;
        if (c >= 0) {
            const int64_t at = isl_find_in_row(rowptr, col, r, c);
            if (at < 0 || at - rowptr[r] >= 0xfffe) err[0] = 1;
            else {
                v = (uint16_t)(at - rowptr[r]);
                if (row_lo[r] > (int32_t)v) atomicMin(row_lo + r, (int32_t)v);
                if (row_hi[r] < (int32_t)v) atomicMax(row_hi + r, (int32_t)v);
            }
        }
        pos[t] = v;
    }
}

template <int G, int U>
__global__ void __launch_bounds__(256) k_gen_gather_rows(const GenGatherParams p) {
    extern __shared__ double gsm[];
    const int sub = threadIdx.x / G, lane = threadIdx.x % G, nsub = blockDim.x / G;
    const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
    double* buf = gsm + (size_t)sub * p.buf_len;
    for (int64_t r = (int64_t)blockIdx.x * nsub + sub; r < p.n_rows; r += (int64_t)gridDim.x * nsub) {
        const int64_t s = p.rowptr[r];
        const int len = (int)(p.rowptr[r + 1] - s);
        const int64_t q0 = p.row_start[r], q1 = p.row_start[r + 1];
        if (q0 == q1 && !p.store) continue;
        int lo = 0, w = len;   // the part of the row this block reaches
        if (p.row_lo) { lo = p.row_lo[r]; w = p.row_hi[r] - lo + 1; if (w <= 0) { lo = 0; w = 0; } }
        for (int k = lane; k < w; k += G) buf[k] = 0.;
        double lift = 0.;
        __syncwarp(mask);
        // the loads of pair q + 1 are in flight while pair q is added into the row buffer
        uint16_t at_n[U]; double v_n[U];
        auto fetch = [&](int64_t q) {
            size_t kr;
            if (p.krow) kr = (size_t)__ldg(p.krow + q);
            else {
                const int32_t pr = __ldg(p.pair + q);
                const int64_t e = pr / p.nr;
                const int i = pr - (int)e * p.nr;
                kr = (size_t)e * p.KR + (p.compact ? i / p.dst : i);
            }
            const double* krow = p.Kbuf + kr * p.KC;
            const uint16_t* pq = p.pos + (size_t)q * p.KC;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int j = lane + G * u;
                const bool in = j < p.KC;
                at_n[u] = in ? __ldg(pq + j) : (uint16_t)0xfffe;
                v_n[u] = in ? __ldg(krow + j) : 0.;
            }
        };
        if (q0 < q1) fetch(q0);
        for (int64_t q = q0; q < q1; q++) {
            uint16_t at[U]; double v[U];
#pragma unroll
            for (int u = 0; u < U; u++) { at[u] = at_n[u]; v[u] = v_n[u]; }
            if (q + 1 < q1) fetch(q + 1);
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (at[u] < 0xfffe) buf[(int)at[u] - lo] += v[u];
                else if (at[u] == 0xffff && v[u] != 0.) {   // column not ACTIVE: Dirichlet lift of a CONSTRAINED DoF, nothing for an inactive one
                    const int32_t pr = __ldg(p.pair + q);   // (rare path: rows next to a Dirichlet boundary)
                    const int64_t e = pr / p.nr; const int i = pr - (int)e * p.nr;
                    const int j = lane + G * u;
                    const int N = p.compact ? j : j / p.dsc, cj = p.compact ? i % p.dst : j % p.dsc;
                    const size_t k = (size_t)p.ed_c[e * p.nc + N] * p.dsc + cj;
                    if (p.st_c[k] == ISL_CONSTRAINED) lift += (p.incremental ? p.presc_c[k] - p.val_c[k] : p.presc_c[k]) * v[u];
                }
            }
            __syncwarp(mask);
        }
        if (p.store) { for (int k = lane; k < len; k += G) { const int kb = k - lo; p.val[s + k] = (kb >= 0 && kb < w) ? buf[kb] : 0.; } }
        else {   // untouched entries (other blocks of the row, other components) are neither read nor written
            for (int k = lane; k < w; k += G) { const double b = buf[k]; if (b != 0.) p.val[s + lo + k] += b; }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) lift += __shfl_xor_sync(mask, lift, o);
        if (lane == 0 && lift != 0.) atomicAdd(p.rhs + r, -lift);
        __syncwarp(mask);
    }
}
