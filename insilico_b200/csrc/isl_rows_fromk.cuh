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
    DevBuf<int32_t> tile_perm;  // launch order of the 128-row tiles of the row kernel: along a Z-curve of their positions
    // producer / consumer pipeline (k_q1hex_pipeline): chunk c = rows [row_b[c], row_b[c+1]) and the elements whose
    // smallest equation lies in that range, sorted positions [elem_b[c], elem_b[c+1]); rows of chunk c only read elements
    // of chunks c - maxback .. c
    bool pipe_ok = false;
    int NC = 0, slots = 0, n_seg = 0, maxback = 0; int64_t chunk_cap = 0, n_items = 0;
    DevBuf<int64_t> d_row_b, d_elem_b, seg_first; DevBuf<int32_t> seg_chunk, e_tiles, r_tiles; DevBuf<uint8_t> seg_kind;
    DevBuf<int> counters;       // [0..1] queue head (64 bit), [2] e_prefix, [3] r_prefix, [4 .. 4+NC) edone, [4+NC .. 4+2NC) rdone
    DevBuf<double> Kring;       // [44][slots * chunk_cap]
};

// how many chunks below its own does a row reach for its elements?
__global__ void k_fromk_maxback(const int32_t* row_pos, int64_t n_rows, const int64_t* row_b, const int64_t* elem_b, int NC, int* maxback) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        int lo = 0, hi = NC;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (row_b[mid] <= r) lo = mid; else hi = mid; }
        const int cr = lo;
        for (int a = 0; a < 8; a++) {
            const int32_t t = row_pos[r * 8 + a];
            if (t < 0) continue;
            int cc = cr;
            while (cc > 0 && (int64_t)t < elem_b[cc]) cc--;
            if ((int64_t)t >= elem_b[cc + 1]) { atomicMax(maxback, 1 << 20); continue; }   // element above its row's chunk: impossible by construction
            atomicMax(maxback, cr - cc);
        }
    }
}

struct FromKParams {
    const double* coords; const int32_t* conn; const int32_t* eorder; int64_t n_elems;
    const int32_t* row_pos; const RowMeta* meta; int64_t n_rows;
    double* K;
    RowsParams r;   // val, rhs, lift tables, flags (patch arrays unused)
    int matrix;     // 0: right-hand side only (body force without a preceding stiffness call)
    int64_t e_lo, e_hi, r_lo, r_hi;   // this launch: sorted element positions [e_lo, e_hi) / rows [r_lo, r_hi)
    const int32_t* tile_perm;         // CTA -> row tile (nullptr: in equation order)
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

// Z-curve key of a row tile: centroid of an element next to its middle row (10 bits per axis inside the bounding box)
__global__ void k_fromk_tile_key(const double* coords, const int32_t* conn, const int32_t* eorder, const int32_t* row_pos, int64_t n_rows,
                                 int nt, int64_t n_tiles, const unsigned long long* mn, const unsigned long long* mx, int32_t* key, int32_t* idx) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = min(t * nt + nt / 2, n_rows - 1);
        int pos = -1;
        for (int a = 0; a < 8 && pos < 0; a++) pos = row_pos[r * 8 + a];
        unsigned int code = 0;
        if (pos >= 0) {
            const int64_t e = eorder[pos];
            for (int d = 0; d < 3; d++) {
                double c = 0.;
                for (int a = 0; a < 8; a++) c += coords[(size_t)conn[e * 8 + a] * 3 + d];
                c *= 0.125;
                const double lo = ord_dbl(mn[d]), hi = ord_dbl(mx[d]);
                const double u = (hi > lo) ? (c - lo) / (hi - lo) : 0.;
                code |= spread3((unsigned int)min(1023.0, max(0.0, u * 1024.0))) << d;
            }
        }
        key[t] = (int32_t)code; idx[t] = (int32_t)t;
    }
}

// element t (sorted position): local matrix by sum factorisation -> K[.][kcol]
template <bool CG>
__device__ __forceinline__ void fk_element(const FromKParams& p, int64_t t, int64_t kcol, int64_t kstride) {
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
    double* out = p.K + kcol;
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = a + 1; b < 8; b++) {
            if (CG) __stcg(out + (size_t)sym_idx(a, b) * kstride, K[sym_idx(a, b)]); else out[(size_t)sym_idx(a, b) * kstride] = K[sym_idx(a, b)];
        }
#pragma unroll
    for (int a = 0; a < 8; a++) {
        if (CG) __stcg(out + (size_t)(36 + a) * kstride, bf[a]); else out[(size_t)(36 + a) * kstride] = bf[a];
    }
}

// the 27 stencil entries and the body-force integral of row r gathered from K; COL maps a sorted element position to its
// column of K.  CG: loads bypass L1 (the producer CTAs of the pipelined kernel write K while this SM may hold stale lines)
template <bool CG, int U = 4, class COL>
__device__ __forceinline__ void fk_gather_row(const FromKParams& p, int64_t r, int64_t kstride, COL&& col, double (&acc)[27], double& body) {
    const int4 q0 = __ldg(reinterpret_cast<const int4*>(p.row_pos + r * 8));
    const int4 q1 = __ldg(reinterpret_cast<const int4*>(p.row_pos + r * 8) + 1);
    const int pos[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    // U slots at a time: 8 U independent coalesced loads are in flight before the first is used (rows without an
    // element in a slot read some valid column with weight zero, so that no load hides behind a branch)
#pragma unroll
    for (int h4 = 0; h4 < 8 / U; h4++) {
        double v[U][8];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int a = h4 * U + u;
            const double* kc = p.K + col(pos[a]);
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const double* src = kc + (size_t)(b == a ? 36 + a : sym_idx(a, b)) * kstride;   // slot b == a: body force
                v[u][b] = CG ? __ldcg(src) : __ldg(src);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int a = h4 * U + u;
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
}

__global__ void __launch_bounds__(128) k_q1hex_elemK(const FromKParams p) {
    const int64_t t = p.e_lo + (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (t >= p.e_hi) return;
    fk_element<false>(p, t, t, p.n_elems);
}

template <int NT, int U = 4, int MINB = 1>
__global__ void __launch_bounds__(NT, MINB) k_q1hex_rows_fromK(const FromKParams p) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* st = smem + (size_t)warp * RG_STAGE;
    // Tiles are launched along a Z-curve of their position in space, not in equation order: the element matrices a tile
    // reads are read again by the tiles of the neighbouring lines and planes, which then run while those lines of K are
    // still in L2 (in equation order the next plane is 66 k rows = 50 MB of traffic away at 256^3)
    const int64_t tile = p.tile_perm ? (int64_t)__ldg(p.tile_perm + blockIdx.x) : (int64_t)blockIdx.x;
    const int64_t r = p.r_lo + tile * NT + tid;
    const bool act = r < p.r_hi;
    RowMeta m;
    double acc[27];
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = 0.;
    if (act) {
        rg_load_meta(m, p.meta + r);
        double body = 0.;
        fk_gather_row<false, U>(p, r, p.n_elems, [](int t) { return (int64_t)max(t, 0); }, acc, body);
        rg_rhs(p.r, m, acc, body);
    }
    if (p.matrix) {
        rg_write_rows(p.r, m, acc, act, st, lane);
        rg_bulk_wait_read();
    }
}

// ---------------------------------------------------------------------------------------------
// The same two stages in ONE persistent kernel, as a producer / consumer pipeline through L2 (k_q1hex_pipeline):
// the element matrices no longer make a round trip through HBM.
//   * work items in one queue (an atomic counter): E(0), E(1), R(0), E(2), R(1), ... where E(c) = tiles of 128 elements
//     of chunk c, R(c) = tiles of 128 rows of chunk c.  A chunk is ~64 k rows and the elements whose smallest equation
//     lies in that range, so rows of chunk c only read elements of chunks <= c.
//   * E tiles write K into a RING of chunk slots (K[44][slots * chunk_cap]), finish with __threadfence + a per-chunk
//     counter; the thread that completes a chunk advances the prefix "all chunks below are complete".  R tiles wait for
//     prefix > c (they were queued behind E(c+1), so the producers they wait for are already running: no deadlock), read
//     K with ld.global.cg (L1 may hold stale lines of a ring slot) and count themselves done; E tiles of chunk c wait
//     until the rows of chunk c - slots + 1 ... have been gathered before they overwrite a ring slot.
//   * ~20 MB of K per chunk, 6 slots: K lives in the 126 MB L2, is overwritten there before it is ever evicted, and
//     FP64-bound producer CTAs share every SM with memory-bound consumer CTAs.
struct PipeParams {
    FromKParams k;
    const int64_t* seg_first;   // [n_seg + 1] first queue item of every segment
    const int32_t* seg_chunk;   // [n_seg] chunk of the segment; kind = E when seg_kind[s] == 0
    const uint8_t* seg_kind;
    const int64_t* row_b; const int64_t* elem_b;   // [NC + 1]
    int n_seg, NC, slots, maxback; int64_t chunk_cap;
    unsigned long long* next;   // queue head
    int* edone; int* rdone;     // [NC] finished E / R tiles per chunk
    int* e_prefix; int* r_prefix;   // chunks [0, prefix) complete
    const int32_t* e_tiles; const int32_t* r_tiles;   // [NC] tiles per chunk
    int64_t n_items;
};

__device__ __forceinline__ void pipe_complete(int* done, const int32_t* tiles, int* prefix, int c, int NC) {
    // called by one thread after the tile's results are visible device-wide
    if (atomicAdd(done + c, 1) + 1 != tiles[c]) return;
    // this chunk is complete: advance the prefix over all complete chunks that follow it
    for (;;) {
        const int pv = atomicAdd(prefix, 0);
        if (pv >= NC) return;
        if (atomicAdd(done + pv, 0) != tiles[pv]) return;
        atomicCAS(prefix, pv, pv + 1);
    }
}
__device__ __forceinline__ void pipe_wait(const int* prefix, int need) {
    while (atomicAdd(const_cast<int*>(prefix), 0) < need) __nanosleep(64);
}

__global__ void __launch_bounds__(128, 3) k_q1hex_pipeline(const PipeParams q) {
    extern __shared__ __align__(16) double smem[];
    __shared__ long long s_item;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* st = smem + (size_t)warp * RG_STAGE;
    const FromKParams& p = q.k;
    const int64_t kstride = (int64_t)q.slots * q.chunk_cap;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = (long long)atomicAdd(q.next, 1ull);
        __syncthreads();
        const int64_t item = s_item;
        if (item >= q.n_items) break;
        int lo = 0, hi = q.n_seg;   // segment of the item
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (q.seg_first[mid] <= item) lo = mid; else hi = mid; }
        const int c = q.seg_chunk[lo];
        const int64_t tile = item - q.seg_first[lo];
        const int64_t slot_base = (int64_t)(c % q.slots) * q.chunk_cap;
        if (q.seg_kind[lo] == 0) {
            // ---- producer: 128 elements of chunk c
            if (c + q.maxback >= q.slots) { if (tid == 0) pipe_wait(q.r_prefix, c + q.maxback - q.slots + 1); __syncthreads(); }   // ring slot free?
            const int64_t t = q.elem_b[c] + tile * 128 + tid;
            if (t < q.elem_b[c + 1]) fk_element<true>(p, t, slot_base + (t - q.elem_b[c]), kstride);
            __threadfence();
            __syncthreads();
            if (tid == 0) pipe_complete(q.edone, q.e_tiles, q.e_prefix, c, q.NC);
        } else {
            // ---- consumer: 128 rows of chunk c
            if (tid == 0) pipe_wait(q.e_prefix, c + 1);
            __syncthreads();
            const int64_t r = q.row_b[c] + tile * 128 + tid;
            const bool act = r < q.row_b[c + 1];
            RowMeta m;
            double acc[27];
#pragma unroll
            for (int k = 0; k < 27; k++) acc[k] = 0.;
            if (act) {
                rg_load_meta(m, p.meta + r);
                double body = 0.;
                const int64_t* eb = q.elem_b; const int NC = q.NC, slots = q.slots; const int64_t cap = q.chunk_cap;
                // sorted position -> (chunk, offset) -> ring column; the element of a row lies in chunk c or just below
                auto col = [&](int t) -> int64_t {
                    if (t < 0) return slot_base;
                    int cc = c;
                    while (cc > 0 && (int64_t)t < eb[cc]) cc--;
                    (void)NC;
                    return (int64_t)(cc % slots) * cap + ((int64_t)t - eb[cc]);
                };
                fk_gather_row<true, 4>(p, r, kstride, col, acc, body);
                rg_rhs(p.r, m, acc, body);
            }
            if (p.matrix) {
                rg_write_rows(p.r, m, acc, act, st, lane);
                rg_bulk_wait_read();
            }
            __syncthreads();
            if (tid == 0) pipe_complete(q.rdone, q.r_tiles, q.r_prefix, c, q.NC);
        }
    }
}
