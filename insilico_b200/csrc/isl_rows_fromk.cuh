// =============================================================================
// isl_rows_fromk.cuh -- Q1-hex scalar Laplace on lattice-like meshes with GENERAL (non-affine) elements, in two
// streaming kernels instead of one patch kernel (included by isl_engine.cu after isl_rowgather.cuh):
//
//   k_q1hex_elemK       one thread per element: the symmetric local matrix by sum factorisation (q1_K_fast, about
//                       1.2 k FP64 instructions) and the body-force integrals, written structure-of-arrays
//                       K[44][n_elems] in a locality-sorted element order.  Every element is computed ONCE (the patch
//                       kernels recompute the halo elements of every patch, 1.5-1.7x), no shared memory, no barrier:
//                       the kernel is bound by the FP64 pipe alone.
//   k_q1hex_rows_fromK  one thread per matrix row in equation order: the (up to) eight elements around the row node
//                       are looked up in a per-row table, the 64 + 8 numbers they contribute are read with coalesced
//                       loads (neighbouring rows read neighbouring elements of the same array), summed into the 27
//                       stencil accumulators and leave through the bulk-copy write-out of isl_rowgather.cuh; a warp's 32
//                       consecutive rows are one contiguous piece of the value array.
//
// Why not fused: the local matrix of a general element is 36 + 8 doubles; a patch of 256 rows needs 140 KB of shared
// memory for its ~400 element instances, which leaves one CTA of 4-8 warps per SM (measured: 7.3-10.6 ms per 256^3
// assembly, FP64 pipe 17 % busy).  Spilling K through L2 / HBM costs 2 x 5.9 GB of traffic but lets both kernels run
// at full occupancy.  Same job as k_q1hex_patch / k_q1hex_rows_general: asmb/StiffnessMatrix.hpp:159-225 +
// asmb/assembleMatrix.hpp:56-130 for heat::Laplace on a scalar Q1 field, fused asmb/BodyForce.hpp:172-205.
// =============================================================================
#pragma once

struct FromKSet {
    bool built = false, ok = false;
    int64_t n_rows = 0, n_elems = 0;
    DevBuf<int32_t> eorder;     // sorted position -> element
    DevBuf<int32_t> row_pos;    // [n_rows][8] sorted position of the element that has the row as local node a, -1: none
    DevBuf<unsigned char> meta; // RowMeta per row (slot[] unused)
    DevBuf<int32_t> lift_nodes;
    DevBuf<double> K;           // [44][n_elems]: 36 symmetric entries (sym_idx), 8 body-force integrals
    // software pipeline: chunk c = rows [row_b[c], row_b[c+1]) and the elements whose smallest equation lies in that
    // range, sorted positions [elem_b[c], elem_b[c+1]); rows of chunk c only need elements of chunks <= c
    std::vector<int64_t> row_b, elem_b;
    std::vector<cudaEvent_t> ev;
    cudaStream_t aux = nullptr; cudaEvent_t ev_start = nullptr;
    ~FromKSet() {
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        if (ev_start) cudaEventDestroy(ev_start);
        if (aux) cudaStreamDestroy(aux);
    }
};

struct FromKParams {
    const double* coords; const int32_t* conn; const int32_t* eorder; int64_t n_elems;
    const int32_t* row_pos; const RowMeta* meta; int64_t n_rows;
    double* K;
    RowsParams r;   // val, rhs, lift tables, flags (patch arrays unused)
    int matrix;     // 0: right-hand side only (body force without a preceding stiffness call)
    int64_t e_lo, e_hi, r_lo, r_hi;   // this launch: sorted element positions [e_lo, e_hi) / rows [r_lo, r_hi)
};

// first sorted position whose key is >= bound[c], for every chunk boundary
__global__ void k_fromk_bounds(const int32_t* sorted_key, int64_t n, const int64_t* bound, int nb, int64_t* out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nb) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)sorted_key[mid] < bound[c]) lo = mid + 1; else hi = mid; }
    out[c] = lo;
}

// locality key of an element: its smallest equation number (rows of one warp then read neighbouring K columns)
__global__ void k_fromk_elem_key(const int32_t* elem_eqn, int64_t n, int32_t* key, int32_t* idx) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int32_t m = 0x7fffffff;
#pragma unroll
        for (int a = 0; a < 8; a++) { const int32_t q = elem_eqn[e * 8 + a]; if (q >= 0 && q < m) m = q; }
        key[e] = m; idx[e] = (int32_t)e;
    }
}

// row -> (sorted position of the) element that has it as local node a; err: a row is local node a of two elements
// (not lattice-like) or an element has the same equation at two local nodes
__global__ void k_fromk_row_pos(const int32_t* eorder, const int32_t* elem_eqn, int64_t n, int32_t* row_pos, int* err) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = eorder[t];
        int32_t q[8];
#pragma unroll
        for (int a = 0; a < 8; a++) q[a] = elem_eqn[e * 8 + a];
#pragma unroll
        for (int a = 0; a < 8; a++) {
            if (q[a] < 0) continue;
            for (int b = 0; b < a; b++) if (q[b] == q[a]) *err = 1;
            if (atomicCAS(row_pos + (size_t)q[a] * 8 + a, -1, (int32_t)t) != -1) *err = 1;
        }
    }
}

__global__ void k_fromk_row_meta(int pass, int64_t n_rows, const int32_t* row_pos, const int32_t* eorder, const int32_t* conn,
                                 const int32_t* node_eqn, const uint8_t* status, const int64_t* rowptr, const int32_t* col,
                                 RowMeta* meta, int32_t* lift_nodes, int* counter, int* err) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_rows; g += (int64_t)gridDim.x * blockDim.x) {
        if (pass == 1 && meta[g].lift < 0) continue;
        int32_t el[8];
#pragma unroll
        for (int a = 0; a < 8; a++) { const int32_t t = row_pos[g * 8 + a]; el[a] = t < 0 ? -1 : eorder[t]; }
        RowMeta m;
        int32_t nbn[27];
        bool cnb = false;
        const bool ok = rg_row_tables_el((int32_t)g, el, conn, node_eqn, status, rowptr, col, m, nbn, cnb);
        if (pass == 0) {
            if (!ok) { *err = 1; continue; }
#pragma unroll
            for (int a = 0; a < 8; a++) m.slot[a] = 0xffff;
            if (cnb) m.lift = atomicAdd(counter, 1);
            meta[g] = m;
        } else {
            int32_t* dst = lift_nodes + (size_t)meta[g].lift * 27;
            for (int k = 0; k < 27; k++) dst[k] = nbn[k];
        }
    }
}

__global__ void __launch_bounds__(128) k_q1hex_elemK(const FromKParams p) {
    const int64_t t = p.e_lo + (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (t >= p.e_hi) return;
    const int64_t e = __ldg(p.eorder + t);
    const int4 c0 = __ldg(reinterpret_cast<const int4*>(p.conn + e * 8));
    const int4 c1 = __ldg(reinterpret_cast<const int4*>(p.conn + e * 8) + 1);
    const int nd[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    double X[8][3];
#pragma unroll
    for (int a = 0; a < 8; a++) {
        const double* c = p.coords + (size_t)nd[a] * 3;
        X[a][0] = __ldg(c); X[a][1] = __ldg(c + 1); X[a][2] = __ldg(c + 2);
    }
    double K[36], detw[8], bf[8];
    q1_K_fast(X, p.r.factor, K, detw, bf, true);
    // the diagonal entries are not stored: the local matrix of the Laplace operator has zero row sums, the row kernel
    // recovers the diagonal of the assembled row from its off-diagonal entries (8 of 44 arrays less to write and read)
    double* out = p.K + t;
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = a + 1; b < 8; b++) out[(size_t)sym_idx(a, b) * p.n_elems] = K[sym_idx(a, b)];
#pragma unroll
    for (int a = 0; a < 8; a++) out[(size_t)(36 + a) * p.n_elems] = bf[a];
}

template <int NT>
__global__ void __launch_bounds__(NT) k_q1hex_rows_fromK(const FromKParams p) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* st = smem + (size_t)warp * RG_STAGE;
    const int64_t r = p.r_lo + (int64_t)blockIdx.x * NT + tid;
    const bool act = r < p.r_hi;
    RowMeta m;
    double acc[27];
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = 0.;
    if (act) {
        rg_load_meta(m, p.meta + r);
        const int4 q0 = __ldg(reinterpret_cast<const int4*>(p.row_pos + r * 8));
        const int4 q1 = __ldg(reinterpret_cast<const int4*>(p.row_pos + r * 8) + 1);
        const int pos[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        double body = 0.;
        // four slots at a time: 32 independent coalesced loads are in flight before the first is used (rows without an
        // element in a slot read column 0 with weight zero, so that no load hides behind a branch)
#pragma unroll
        for (int h4 = 0; h4 < 2; h4++) {
            double v[4][8];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int a = h4 * 4 + u;
                const double* kc = p.K + max(pos[a], 0);
#pragma unroll
                for (int b = 0; b < 8; b++) v[u][b] = __ldg(kc + (size_t)(b == a ? 36 + a : sym_idx(a, b)) * p.n_elems);   // slot b == a: body force
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int a = h4 * 4 + u;
                const double wgt = pos[a] >= 0 ? 1.0 : 0.0;
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    if (b == a) body = fma(wgt, v[u][b], body);
                    else acc[rg_kidx(a, b)] = fma(wgt, v[u][b], acc[rg_kidx(a, b)]);
                }
            }
        }
        // diagonal of the row: minus the sum of the 26 other stencil entries (zero row sums of every local matrix)
        double dsum = 0.;
#pragma unroll
        for (int k = 0; k < 27; k++) if (k != 13) dsum += acc[k];
        acc[13] = -dsum;
        rg_rhs(p.r, m, acc, body);
    }
    if (p.matrix) {
        rg_write_rows(p.r, m, acc, act, st, lane);
        rg_bulk_wait_read();
    }
}
