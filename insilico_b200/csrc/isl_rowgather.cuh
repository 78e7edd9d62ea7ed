// =============================================================================
// isl_rowgather.cuh -- row-gather assembly of the Q1-hex scalar Laplace matrix on all-affine, lattice-like meshes.
// (included by isl_engine.cu inside its anonymous namespace after isl_patch.cuh; the per-thread routines are also
// compiled for the host by tests/emu/rowgather_emu.cpp, which checks them against the oracle without a GPU)
//
// Same job as k_q1hex_patch_affine (asmb/StiffnessMatrix.hpp:159-225 + asmb/assembleMatrix.hpp:56-130 for
// heat::Laplace on a scalar Q1 field, fused asmb/BodyForce.hpp:172-205), but turned around: instead of every element
// ADDING its 8x8 matrix into shared-memory rows (64 read-modify-writes per element in eight barrier-separated phases,
// halo elements recomputed), every owned ROW gathers its entries:
//
//   phase 1, one thread per element instance of the patch:  D_c = (kappa w / det J) (cof^T cof)_c, c = 0..5, and
//            det J w, from the constant Jacobian of the affine element -> shared memory (7 doubles per instance);
//   phase 2, one thread per owned row i:  the row node is local node a of at most one element e_a (a = 0..7, the
//            "lattice property" already required by the patch kernels).  A neighbour j that is local node b of e_a sits
//            at the lattice offset  d = bits(b) - bits(a) in {-1,0,1}^3  from i, so
//                 A[i, j(d)] = sum_{(a,b): bits(b)-bits(a) = d}  sum_c D_c(e_a) C_c[a][b]
//            with compile-time (a, b, d): 64 x 6 DFMA with constant-bank operands into 27 register accumulators.
//            No shared-memory accumulator, no barriers, no recomputation of the matrix part for halo elements.
//   write-out: the 27 values go through a small per-warp staging buffer into CSR order (uint8 position per stencil
//            neighbour, computed once on the device) and leave as row-contiguous stores; every entry written once.
//   Dirichlet lift: a stencil neighbour without a CSR position that is CONSTRAINED contributes g_j A[i,j] to the
//            right-hand side (rows next to such nodes carry an index into a table of neighbour node ids).
//
// Eligibility (checked per row when the tables are built, otherwise the patch kernels are used): lattice property,
// the same neighbour reached through different elements gets the same offset, every CSR entry of the row is a
// stencil neighbour or the row is marked partial and zero-filled first (so a complete row is written and the
// lazily-zeroed matrix needs no memset).
// =============================================================================
#pragma once

#ifdef __CUDACC__
#define RG_HD __host__ __device__ __forceinline__
#else
#define RG_HD inline
#endif

// hierarchic vertex a of the hexahedron -> (xi, eta, zeta) bits (base/mesh/HierarchicOrder.hpp:200-272:
// 0 (0,0,0) 1 (1,0,0) 2 (1,1,0) 3 (0,1,0) 4 (0,0,1) 5 (1,0,1) 6 (1,1,1) 7 (0,1,1))
RG_HD constexpr int rg_bx(int a) { return (a == 1 || a == 2 || a == 5 || a == 6) ? 1 : 0; }
RG_HD constexpr int rg_by(int a) { return (a == 2 || a == 3 || a == 6 || a == 7) ? 1 : 0; }
RG_HD constexpr int rg_bz(int a) { return a >= 4 ? 1 : 0; }
// stencil index of the neighbour that is local node b of the element in which the row node is local node a
RG_HD constexpr int rg_kidx(int a, int b) {
    return (rg_bz(b) - rg_bz(a) + 1) * 9 + (rg_by(b) - rg_by(a) + 1) * 3 + (rg_bx(b) - rg_bx(a) + 1);
}

struct alignas(16) RowMeta {
    uint16_t slot[8];  // patch-local element instance that has this row as local node a (0xffff: none)
    uint8_t pos[27];   // position of stencil neighbour k inside the CSR row (0xff: no entry)
    uint8_t nnz;       // entries of the row (bits 0-6); bit 7: partial row, has entries that are no stencil neighbours
    int32_t lift;      // row next to CONSTRAINED nodes: index into the table of neighbour node ids, else -1
    int32_t grow;      // global row (equation number)
    int32_t pad;
    int64_t rowstart;  // rowptr[grow]
};
static_assert(sizeof(RowMeta) == 64, "RowMeta is read as four 16-byte words");

// tables: device = the engine's constant memory, host emulation = plain arrays filled by the test harness
#ifdef __CUDA_ARCH__
#define RG_NSUM(i) c_q1_Nsum[i]
#else
static double rg_host_nsum[8];
#define RG_NSUM(i) rg_host_nsum[i]
#endif

// phase 1: the six numbers E_c = (kappa / det J) (cof^T cof)_c, c = xx yy zz xy xz yz, and det J w of an affine
// element from its hierarchic nodes 0, 1 (xi), 3 (eta), 4 (zeta)
RG_HD void rg_instance(const double* x0, const double* x1, const double* x3, const double* x4, double factor, double w,
                       double (&E)[6], double& dw) {
    double J[3][3];
    for (int d = 0; d < 3; d++) { J[d][0] = x1[d] - x0[d]; J[d][1] = x3[d] - x0[d]; J[d][2] = x4[d] - x0[d]; }
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
    const double s = factor / det;
    dw = det * w;
    E[0] = s * (co[0][0] * co[0][0] + co[1][0] * co[1][0] + co[2][0] * co[2][0]);
    E[1] = s * (co[0][1] * co[0][1] + co[1][1] * co[1][1] + co[2][1] * co[2][1]);
    E[2] = s * (co[0][2] * co[0][2] + co[1][2] * co[1][2] + co[2][2] * co[2][2]);
    E[3] = s * (co[0][0] * co[0][1] + co[1][0] * co[1][1] + co[2][0] * co[2][1]);
    E[4] = s * (co[0][0] * co[0][2] + co[1][0] * co[1][2] + co[2][0] * co[2][2]);
    E[5] = s * (co[0][1] * co[0][2] + co[1][1] * co[1][2] + co[2][1] * co[2][2]);
}

// phase 2 for affine elements: the 27 stencil entries of one row from the E_c of the (up to) eight elements around its
// node.  On the reference cube the trilinear integrals are exact rationals (the 2x2x2 Gauss rule of the reference
// integrates them exactly, so this equals its quadrature up to rounding):
//   int d_x N_a d_x N_b = s_x(a) s_x(b) m_y m_z,   m = 1/3 (a, b on the same side of that axis) or 1/6,  s = -1 / +1,
//   int (d_x N_a d_y N_b + d_y N_a d_x N_b) = +-(1/2) m_z s_x(a) s_y(a)  when a, b agree on both or differ on both of
//                                              x, y (+ / -), else 0.
// For a FIXED stencil offset d = bits(b) - bits(a) these factors do not depend on the element any more, only on d (and,
// for the mixed terms, on the signs s(a) of the row node inside the element), so
//   A[i, i+d] = sum_c coef_c(d) * S_c(d),   S_c(d) = (signed) sum of E_c over the elements that contain both nodes.
// The 27 sums S_c(d) are a tensor product of {element on the - side, both, element on the + side} per axis and are
// formed by a three-level tree: 19 additions per diagonal component, 11 per mixed component (only the offsets with
// d_p = d_q = 0 or both non-zero are needed), then one multiply-add per non-zero coefficient: about 215 FP64
// operations per row and ten distinct constants, instead of 448 operations and 216 constant-table entries when the
// eight 8-entry local rows are formed one by one.
//   ld(A, c): E_c (c < 6) or det J w (c = 6) of the element that has the row as local node A, 0 when there is none.
RG_HD constexpr int rg_bit(int axis, int a) { return axis == 0 ? rg_bx(a) : axis == 1 ? rg_by(a) : rg_bz(a); }
// hierarchic vertex with the given (x, y, z) bits
RG_HD constexpr int rg_vertex(int bx, int by, int bz) {
    for (int a = 0; a < 8; a++)
        if (rg_bx(a) == bx && rg_by(a) == by && rg_bz(a) == bz) return a;
    return -1;
}
// one tree level along an axis: t = 0 (offset -1: the row node is on the element's + side, bit 1), t = 1 (offset 0:
// both elements), t = 2 (offset +1: bit 0).  sgn: weight the elements with the sign s(a) = -1 (bit 0) / +1 (bit 1).
RG_HD void rg_level(double e0, double e1, bool sgn, double& tm, double& t0, double& tp) {
    tm = e1;
    tp = sgn ? -e0 : e0;
    t0 = sgn ? e1 - e0 : e0 + e1;
}
template <class LD>
RG_HD void rg_row_affine(LD&& ld, double (&acc)[27], double& body) {
    constexpr double mu[3] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 6.0};   // by offset index t = d + 1
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int c = 0; c < 6; c++) {
        // axes: diagonal component c < 3 -> (c, c); mixed 3 -> (0,1), 4 -> (0,2), 5 -> (1,2)
        const int p = (c < 3) ? c : (c == 5 ? 1 : 0), q = (c < 3) ? c : (c == 3 ? 1 : 2);
        const bool mixed = c >= 3;
        const bool sg[3] = {mixed && (p == 0 || q == 0), mixed && (p == 1 || q == 1), mixed && (p == 2 || q == 2)};
        double X[2][2][3];   // [bz][by][tx]
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int bz = 0; bz < 2; bz++)
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int by = 0; by < 2; by++)
                rg_level(ld(rg_vertex(0, by, bz), c), ld(rg_vertex(1, by, bz), c), sg[0], X[bz][by][0], X[bz][by][1], X[bz][by][2]);
        double Y[2][3][3];   // [bz][ty][tx]
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int bz = 0; bz < 2; bz++)
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int tx = 0; tx < 3; tx++)
                rg_level(X[bz][0][tx], X[bz][1][tx], sg[1], Y[bz][0][tx], Y[bz][1][tx], Y[bz][2][tx]);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int ty = 0; ty < 3; ty++)
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int tx = 0; tx < 3; tx++) {
                double Z[3];
                rg_level(Y[0][ty][tx], Y[1][ty][tx], sg[2], Z[0], Z[1], Z[2]);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
                for (int tz = 0; tz < 3; tz++) {
                    const int t[3] = {tx, ty, tz};
                    double coef;
                    if (!mixed) {
                        const int o1 = (c + 1) % 3, o2 = (c + 2) % 3;
                        coef = (t[c] == 1 ? 1.0 : -1.0) * mu[t[o1]] * mu[t[o2]];
                    } else {
                        const int r = 3 - p - q;
                        const bool zp = t[p] == 1, zq = t[q] == 1;
                        if (zp != zq) continue;   // the mixed integrals vanish when a, b agree on exactly one of the two axes
                        coef = (zp ? 0.5 : -0.5) * mu[t[r]];
                    }
                    acc[tz * 9 + ty * 3 + tx] = fma(coef, Z[tz], acc[tz * 9 + ty * 3 + tx]);
                }
            }
    }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int a = 0; a < 8; a++) body = fma(ld(a, 6), RG_NSUM(a), body);
}

// general (non-affine) elements: the element's symmetric local matrix (36 numbers) and its body-force integrals
// sum_q N_a w det J (8 numbers) are taken from the patch's instance table instead of being formed from D
template <int A, class LoadK>
RG_HD void rg_add_slot_general(LoadK&& ld, double (&acc)[27], double& body) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int b = 0; b < 8; b++) acc[rg_kidx(A, b)] += ld(sym_idx(A, b));
    body += ld(36 + A);
}

// binary search of column c in CSR row r; -1 when absent
RG_HD int64_t rg_find(const int64_t* rowptr, const int32_t* col, int32_t r, int32_t c) {
    int64_t lo = rowptr[r], hi = rowptr[r + 1];
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (col[mid] < c) lo = mid + 1; else hi = mid; }
    return (lo < rowptr[r + 1] && col[lo] == c) ? lo : -1;
}

// tables of one owned row: stencil neighbours through the (up to) eight elements around the row node, their CSR
// positions, whether a neighbour is CONSTRAINED.  el[a] = element that has the row node as local node a (-1: none).
// Returns false when the row is not eligible (see header).  nbn[k] = global node of stencil neighbour k or -1.
RG_HD bool rg_row_tables_el(int32_t g, const int32_t (&el)[8], const int32_t* conn,
                            const int32_t* node_eqn, const uint8_t* status, const int64_t* rowptr, const int32_t* col,
                            RowMeta& m, int32_t (&nbn)[27], bool& constrained_nb) {
    for (int k = 0; k < 27; k++) { nbn[k] = -1; m.pos[k] = 0xff; }
    for (int a = 0; a < 8; a++) {
        if (el[a] < 0) continue;
        const int32_t* ce = conn + (size_t)el[a] * 8;
        for (int b = 0; b < 8; b++) {
            const int k = rg_kidx(a, b);
            if (nbn[k] >= 0 && nbn[k] != ce[b]) return false;  // two elements disagree about the neighbour at this offset
            nbn[k] = ce[b];
        }
    }
    int cnt = 0;
    constrained_nb = false;
    for (int k = 0; k < 27; k++) {
        if (nbn[k] < 0) continue;
        for (int k2 = 0; k2 < k; k2++) if (nbn[k2] == nbn[k]) return false;  // one node at two offsets (degenerate connectivity)
        const int32_t cq = node_eqn[nbn[k]];
        if (cq >= 0) {
            const int64_t s = rg_find(rowptr, col, g, cq);
            if (s < 0) return false;
            m.pos[k] = (uint8_t)(s - rowptr[g]);
            cnt++;
        } else if (status[nbn[k]] == ISL_CONSTRAINED) constrained_nb = true;
    }
    // A row may have MORE entries than stencil neighbours found through the owned elements (columns that only the
    // pattern-only halo elements of another partition contribute, isl_mesh_set_owned): such a row is "partial", the
    // kernel zero-fills it before it scatters.  More than 27 entries do not fit the staging buffer.
    const int64_t rn = rowptr[g + 1] - rowptr[g];
    if ((int64_t)cnt > rn || rn > 27) return false;
    m.nnz = (uint8_t)(rn | (cnt < rn ? 0x80 : 0));
    m.lift = -1;
    m.grow = g; m.pad = 0; m.rowstart = rowptr[g];
    return true;
}
// the same for a row of a patch: slot[a] = patch-local element instance (0xffff: none), inst_elem = the patch's instance list
RG_HD bool rg_row_tables(int32_t g, const uint16_t* slot, const int32_t* inst_elem, const int32_t* conn,
                         const int32_t* node_eqn, const uint8_t* status, const int64_t* rowptr, const int32_t* col,
                         RowMeta& m, int32_t (&nbn)[27], bool& constrained_nb) {
    int32_t el[8];
    for (int a = 0; a < 8; a++) { m.slot[a] = slot[a]; el[a] = (slot[a] == 0xffff) ? -1 : inst_elem[slot[a]]; }
    return rg_row_tables_el(g, el, conn, node_eqn, status, rowptr, col, m, nbn, constrained_nb);
}

// ---------------------------------------------------------------------------------------------
// write-out of a warp's (up to) 32 rows, shared by the device kernels and the host emulation.
// The rows of a patch are sorted by equation number, so consecutive lanes often hold consecutive CSR rows; a maximal
// piece of consecutive rows inside the warp is a SEGMENT: one contiguous piece of the value array.  The rows are put
// into the warp's staging buffer in CSR order (position bytes of RowMeta), segment after segment; the start of every
// segment is shifted by 0 or 1 doubles so that staging address and global address have the same 16-byte phase, which is
// what the bulk copy (cp.async.bulk shared -> global, SASS UBLKCP) needs.  Offsets of lane l:
//   off  = exclusive prefix sum of the row lengths,  seg = number of segment heads at lanes <= l, minus one,
//   fix  = (rowstart(head) ^ (off(head) + 2 seg)) & 1,  staging offset = off + 2 seg + fix.
constexpr int RG_STAGE = 32 * 27 + 2 * 32 + 2;   // doubles per warp

struct RgLane { int n, off, seg, fix, head_lane, next_head; bool head; };

// lock-step evaluation for one warp on the host (the device kernel does the same with shuffles and ballots)
inline void rg_segments_host(const int (&n)[32], const int64_t (&rs)[32], RgLane (&L)[32], int& total) {
    int run = 0; uint32_t headmask = 0;
    for (int l = 0; l < 32; l++) {
        L[l].n = n[l]; L[l].off = run; run += n[l];
        L[l].head = n[l] > 0 && (l == 0 || rs[l] != rs[l - 1] + n[l - 1]);
        if (L[l].head) headmask |= 1u << l;
    }
    total = run;
    for (int l = 0; l < 32; l++) {
        const uint32_t le = headmask & (0xffffffffu >> (31 - l));
        L[l].seg = __builtin_popcount(le) - 1;
        L[l].head_lane = le ? 31 - __builtin_clz(le) : 0;
        const uint32_t above = (l < 31) ? (headmask >> (l + 1)) : 0u;
        L[l].next_head = above ? l + 1 + __builtin_ctz(above) : 32;
    }
    for (int l = 0; l < 32; l++) {
        const int hl = L[l].head_lane;
        L[l].fix = (int)((rs[hl] ^ (int64_t)(L[hl].off + 2 * L[l].seg)) & 1);
    }
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device side
struct RowsParams {
    const double* coords;
    const int32_t* p_inst_off; const int32_t* p_row_off; const int32_t* p_node_off;
    const int32_t* nodes;      // global node id of every local node
    const uint16_t* i_lnode;   // [inst][8] local node index
    const RowMeta* meta;       // per owned row, patch after patch
    const int32_t* lift_nodes; // [flagged rows][27] neighbour node ids
    const uint8_t* status; const double* presc; const double* values;
    double* val; double* rhs;
    double factor; int incremental; int store_mode;
    int body; double f0;
    int node_cap, inst_cap, n_patches, resident;
    int matrix;   // 0: right-hand side only (body force without a stiffness call before it)
    const int32_t* perm; int patch_base;   // launch order: patch = perm[blockIdx.x + patch_base] (perm may be null)
};

// preprocessing: one CTA per patch, threads over its rows.  pass 0 fills meta and numbers the rows that
// have CONSTRAINED neighbours (counter[0]); pass 1 writes their neighbour node ids.  err[0] != 0: some row is not eligible.
__global__ void k_row_meta(int pass, const int32_t* p_row_off, const int32_t* p_inst_off, const int32_t* rows,
                           const uint16_t* rslot, const int32_t* inst_elem, const int32_t* conn, const int32_t* node_eqn,
                           const uint8_t* status, const int64_t* rowptr, const int32_t* col, RowMeta* meta,
                           int32_t* lift_nodes, int* counter, int* err) {
    const int pid = blockIdx.x;
    const int r0 = p_row_off[pid], nrows = p_row_off[pid + 1] - r0;
    const int32_t* ie = inst_elem + p_inst_off[pid];
    for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
        const int32_t g = rows[r0 + r];
        if (pass == 0) {
            RowMeta m;
            int32_t nbn[27];
            bool cnb = false;
            if (!rg_row_tables(g, rslot + (size_t)(r0 + r) * 8, ie, conn, node_eqn, status, rowptr, col, m, nbn, cnb)) { *err = 1; continue; }
            if (cnb) m.lift = atomicAdd(counter, 1);
            meta[r0 + r] = m;
        } else if (meta[r0 + r].lift >= 0) {
            RowMeta m;
            int32_t nbn[27];
            bool cnb = false;
            rg_row_tables(g, rslot + (size_t)(r0 + r) * 8, ie, conn, node_eqn, status, rowptr, col, m, nbn, cnb);
            int32_t* dst = lift_nodes + (size_t)meta[r0 + r].lift * 27;
            for (int k = 0; k < 27; k++) dst[k] = nbn[k];
        }
    }
}

__device__ __forceinline__ void rg_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// matrix write-out of a warp's 32 rows (see the comment at RG_STAGE): staging in CSR order, then one bulk copy
// (store mode) or bulk reduction (accumulate mode: cp.reduce.async.bulk .add.f64, SASS UBLKRED) per segment, issued
// by the segment's head lane; an odd first / last element of a segment goes out as a plain store / atomic add.
__device__ __forceinline__ void rg_write_rows(const RowsParams& p, const RowMeta& m, const double (&acc)[27], bool act, double* st,
                                              int lane) {
    const int n = act ? (m.nnz & 0x7f) : 0;
    const int64_t rs = act ? m.rowstart : 0;
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const int off = incl - n;
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int64_t rs_prev = __shfl_up_sync(0xffffffffu, rs, 1);
    const int n_prev = __shfl_up_sync(0xffffffffu, n, 1);
    const bool head = n > 0 && (lane == 0 || rs != rs_prev + n_prev);
    const uint32_t headmask = __ballot_sync(0xffffffffu, head);
    const uint32_t le = headmask & (0xffffffffu >> (31 - lane));
    const int seg = __popc(le) - 1;
    const int head_lane = le ? 31 - __clz(le) : 0;
    const uint32_t above = (lane < 31) ? (headmask >> (lane + 1)) : 0u;
    const int next_head = above ? lane + __ffs(above) : 32;
    const int head_off = __shfl_sync(0xffffffffu, off, head_lane);
    const int64_t head_rs = __shfl_sync(0xffffffffu, rs, head_lane);
    const int fix = (int)((head_rs ^ (int64_t)(head_off + 2 * seg)) & 1);
    const int so = off + 2 * seg + fix;
    const int end_off = __shfl_sync(0xffffffffu, off, next_head & 31);
    const int seg_len = (next_head < 32 ? end_off : total) - off;   // meaningful on head lanes
    if (n > 0) {
        if (m.nnz & 0x80) {
#pragma unroll
            for (int k = 0; k < 27; k++)
                if (k < n) st[so + k] = 0.;
        }
#pragma unroll
        for (int k = 0; k < 27; k++)
            if (m.pos[k] != 0xff) st[so + m.pos[k]] = acc[k];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (head) {
        int64_t g0 = rs;
        int s0 = so, len = seg_len;
        if (g0 & 1) {
            if (p.store_mode) p.val[g0] = st[s0]; else atomicAdd(p.val + g0, st[s0]);
            g0++; s0++; len--;
        }
        if (len & 1) {
            if (p.store_mode) p.val[g0 + len - 1] = st[s0 + len - 1]; else atomicAdd(p.val + g0 + len - 1, st[s0 + len - 1]);
            len--;
        }
        if (len > 0) {
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(st + s0);
            if (p.store_mode)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.val + g0), "r"(sa), "r"(len * 8) : "memory");
            else
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(p.val + g0), "r"(sa), "r"(len * 8) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}

// right-hand side of one row: Dirichlet lift of CONSTRAINED stencil neighbours (assembleMatrix.hpp:56-130) and body
// force; added with a reduction (no round trip: every row is owned by exactly one thread of one CTA)
__device__ __forceinline__ void rg_rhs(const RowsParams& p, const RowMeta& m, const double (&acc)[27], double body) {
    double lift = 0.;
    if (m.lift >= 0 && p.matrix) {
        const int32_t* ln = p.lift_nodes + (size_t)m.lift * 27;
#pragma unroll
        for (int k = 0; k < 27; k++) {
            const int32_t nd = __ldg(ln + k);
            if (nd >= 0 && m.pos[k] == 0xff && p.status[nd] == ISL_CONSTRAINED) {
                const double gv = p.incremental ? p.presc[nd] - p.values[nd] : p.presc[nd];
                lift = fma(gv, acc[k], lift);
            }
        }
    }
    const double v = (p.body ? p.f0 * body : 0.) - lift;
    if (v != 0.) atomicAdd(p.rhs + m.grow, v);
}

__device__ __forceinline__ void rg_load_meta(RowMeta& m, const RowMeta* src) {
    const int4* mp = reinterpret_cast<const int4*>(src);
    int4* md = reinterpret_cast<int4*>(&m);
    md[0] = __ldg(mp); md[1] = __ldg(mp + 1); md[2] = __ldg(mp + 2); md[3] = __ldg(mp + 3);
}

// the patch that will run on this CTA's slot next is about `resident` patches ahead: pull its row tables, instance
// nodes and node list into L2 (cp.async.bulk.prefetch.L2, SASS UBLKPF) so that its prologue hits L2 instead of DRAM
__device__ __forceinline__ void rg_prefetch_patch(const RowsParams& p, int pid) {
    if (pid >= p.n_patches) return;
    const int r0 = p.p_row_off[pid], r1 = p.p_row_off[pid + 1];
    const int e0 = p.p_inst_off[pid], e1 = p.p_inst_off[pid + 1];
    const int n0 = p.p_node_off[pid], n1 = p.p_node_off[pid + 1];
    const char* b0 = reinterpret_cast<const char*>(p.meta + r0);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(b0), "r"((r1 - r0) * 64) : "memory");
    const char* b1 = reinterpret_cast<const char*>(p.i_lnode + (size_t)e0 * 8);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(b1), "r"((e1 - e0) * 16) : "memory");
    const size_t a2 = reinterpret_cast<size_t>(p.nodes + n0) & ~(size_t)15;
    const int nb2 = (int)((reinterpret_cast<size_t>(p.nodes + n1) + 15 - a2) & ~(size_t)15);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a2), "r"(nb2) : "memory");
}

// nodal coordinates of the patch -> shared memory (two dependent global loads: node id, then its coordinates)
template <int NT>
__device__ __forceinline__ void rg_load_coords(const RowsParams& p, double* sX, int n0, int nnodes, int tid) {
    constexpr int U = 4;
    for (int nb = 0; nb < nnodes; nb += U * NT) {
        int32_t g[U];
#pragma unroll
        for (int i = 0; i < U; i++) { const int n = nb + i * NT + tid; g[i] = (n < nnodes) ? __ldg(p.nodes + n0 + n) : 0; }
        double x[U][3];
#pragma unroll
        for (int i = 0; i < U; i++) {
            const double* c = p.coords + (size_t)g[i] * 3;
            x[i][0] = __ldg(c); x[i][1] = __ldg(c + 1); x[i][2] = __ldg(c + 2);
        }
#pragma unroll
        for (int i = 0; i < U; i++) {
            const int n = nb + i * NT + tid;
            if (n < nnodes) { sX[n * 3] = x[i][0]; sX[n * 3 + 1] = x[i][1]; sX[n * 3 + 2] = x[i][2]; }
        }
    }
}

// one CTA per patch.  Shared memory: sD[7][inst_cap] | (sX[node_cap][3]  aliased after phase 1 by  stage[NT/32][RG_STAGE]);
// instance inst_cap-1 is a zero element (rows with fewer than eight elements around them)
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_q1hex_rows_affine(const RowsParams p) {
    extern __shared__ __align__(16) double smem[];
    double* sD = smem;
    double* sX = sD + (size_t)7 * p.inst_cap;
    double* stage = sX;
    const int tid = threadIdx.x;
    const int slot_id = blockIdx.x + p.patch_base;
    const int pid = p.perm ? p.perm[slot_id] : slot_id;
    const int r0 = p.p_row_off[pid], nrows = p.p_row_off[pid + 1] - r0;
    const int n0 = p.p_node_off[pid], nnodes = p.p_node_off[pid + 1] - n0;
    const int e0 = p.p_inst_off[pid], ninst = p.p_inst_off[pid + 1] - e0;
    // everything that does not depend on other threads is requested first: instance nodes, the row table of the first row
    constexpr int UI = 4;
    int4 ln4[UI];
#pragma unroll
    for (int u = 0; u < UI; u++) {
        const int i = u * NT + tid;
        ln4[u] = (i < ninst) ? __ldg(reinterpret_cast<const int4*>(p.i_lnode + (size_t)(e0 + i) * 8)) : make_int4(0, 0, 0, 0);
    }
    RowMeta m;
    if (tid < nrows) rg_load_meta(m, p.meta + r0 + tid);
    rg_load_coords<NT>(p, sX, n0, nnodes, tid);
    if (tid < 7) sD[tid * p.inst_cap + p.inst_cap - 1] = 0.;
    if (tid == NT - 1 && slot_id + p.resident < p.n_patches) rg_prefetch_patch(p, p.perm ? p.perm[slot_id + p.resident] : slot_id + p.resident);
    __syncthreads();
    // phase 1: element instances
    const double w = c_q1_w[0];
    auto instance = [&](int i, const int4& l4) {
        const int n0l = l4.x & 0xffff, n1l = (unsigned)l4.x >> 16, n3l = (unsigned)l4.y >> 16, n4l = l4.z & 0xffff;
        double E[6], dw;
        rg_instance(sX + n0l * 3, sX + n1l * 3, sX + n3l * 3, sX + n4l * 3, p.factor, w, E, dw);
#pragma unroll
        for (int c = 0; c < 6; c++) sD[c * p.inst_cap + i] = E[c];
        sD[6 * p.inst_cap + i] = dw;
    };
#pragma unroll
    for (int u = 0; u < UI; u++) {
        const int i = u * NT + tid;
        if (i < ninst) instance(i, ln4[u]);
    }
    for (int i = UI * NT + tid; i < ninst; i += NT)
        instance(i, __ldg(reinterpret_cast<const int4*>(p.i_lnode + (size_t)(e0 + i) * 8)));
    __syncthreads();  // sD complete; sX is dead from here on (stage aliases it)
    // phase 2: owned rows
    const int lane = tid & 31, warp = tid >> 5;
    double* st = stage + (size_t)warp * RG_STAGE;
    const int zs = p.inst_cap - 1;
    for (int rb = 0; rb < nrows; rb += NT) {
        const int r = rb + tid;
        const bool act = r < nrows;
        if (rb > 0) {
            if (act) rg_load_meta(m, p.meta + r0 + r);
            rg_bulk_wait_read();   // the bulk copies of the previous round have read the staging buffer
            __syncwarp();
        }
        if (r + NT < nrows) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.meta + r0 + r + NT));
        double acc[27];
#pragma unroll
        for (int k = 0; k < 27; k++) acc[k] = 0.;
        if (act) {
            int sl[8];
#pragma unroll
            for (int a = 0; a < 8; a++) sl[a] = min((int)m.slot[a], zs);
            double body = 0.;
            rg_row_affine([&](int a, int c) { return sD[c * p.inst_cap + sl[a]]; }, acc, body);
            rg_rhs(p, m, acc, body);
        }
        if (p.matrix) rg_write_rows(p, m, acc, act, st, lane);
    }
    rg_bulk_wait_read();   // shared memory must stay valid until the bulk copies have read it
}

// ---------------------------------------------------------------------------------------------
// Same row-gather scheme for general (non-affine) elements: phase 1 computes the symmetric local matrix with the
// sum-factorised q1_K_fast (isl_patch.cuh) and keeps it in shared memory, 44 doubles per element instance
// (sK[44][inst_cap]); phase 2 only adds.  Shared memory: sK | (sX aliased after phase 1 by the staging buffer).
template <int A>
__device__ __forceinline__ void rg_gather_slot_general(const RowMeta& m, const double* sK, int cap, double (&acc)[27], double& body) {
    const int s = m.slot[A];
    if (s != 0xffff) rg_add_slot_general<A>([&](int i) { return sK[i * cap + s]; }, acc, body);
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_q1hex_rows_general(const RowsParams p) {
    extern __shared__ __align__(16) double smem[];
    double* sK = smem;
    double* sX = sK + (size_t)44 * p.inst_cap;
    double* stage = sX;
    const int tid = threadIdx.x;
    const int slot_id = blockIdx.x + p.patch_base;
    const int pid = p.perm ? p.perm[slot_id] : slot_id;
    const int r0 = p.p_row_off[pid], nrows = p.p_row_off[pid + 1] - r0;
    const int n0 = p.p_node_off[pid], nnodes = p.p_node_off[pid + 1] - n0;
    const int e0 = p.p_inst_off[pid], ninst = p.p_inst_off[pid + 1] - e0;
    constexpr int UI = 2;
    int4 ln4[UI];
#pragma unroll
    for (int u = 0; u < UI; u++) {
        const int i = u * NT + tid;
        ln4[u] = (i < ninst) ? __ldg(reinterpret_cast<const int4*>(p.i_lnode + (size_t)(e0 + i) * 8)) : make_int4(0, 0, 0, 0);
    }
    rg_load_coords<NT>(p, sX, n0, nnodes, tid);
    if (tid == NT - 1 && slot_id + p.resident < p.n_patches) rg_prefetch_patch(p, p.perm ? p.perm[slot_id + p.resident] : slot_id + p.resident);
    __syncthreads();
    // phase 1: local matrices of the element instances
    auto instance = [&](int i, const int4& l4) {
        const int ln[8] = {l4.x & 0xffff, (int)((unsigned)l4.x >> 16), l4.y & 0xffff, (int)((unsigned)l4.y >> 16),
                           l4.z & 0xffff, (int)((unsigned)l4.z >> 16), l4.w & 0xffff, (int)((unsigned)l4.w >> 16)};
        double X[8][3];
#pragma unroll
        for (int a = 0; a < 8; a++) { X[a][0] = sX[ln[a] * 3]; X[a][1] = sX[ln[a] * 3 + 1]; X[a][2] = sX[ln[a] * 3 + 2]; }
        double K[36], detw[8], bf[8];
        q1_K_fast(X, p.factor, K, detw, bf, true);
#pragma unroll
        for (int k = 0; k < 36; k++) sK[k * p.inst_cap + i] = K[k];
#pragma unroll
        for (int a = 0; a < 8; a++) sK[(36 + a) * p.inst_cap + i] = bf[a];
    };
#pragma unroll 1
    for (int u = 0; u < UI; u++) {
        const int i = u * NT + tid;
        if (i < ninst) instance(i, ln4[u]);
    }
#pragma unroll 1
    for (int i = UI * NT + tid; i < ninst; i += NT)
        instance(i, __ldg(reinterpret_cast<const int4*>(p.i_lnode + (size_t)(e0 + i) * 8)));
    __syncthreads();  // sK complete; sX is dead from here on (stage aliases it)
    // phase 2: owned rows (identical to k_q1hex_rows_affine apart from the slot routine)
    const int lane = tid & 31, warp = tid >> 5;
    double* st = stage + (size_t)warp * RG_STAGE;
    for (int rb = 0; rb < nrows; rb += NT) {
        const int r = rb + tid;
        const bool act = r < nrows;
        RowMeta m;
        if (act) rg_load_meta(m, p.meta + r0 + r);
        if (rb > 0) { rg_bulk_wait_read(); __syncwarp(); }
        double acc[27];
#pragma unroll
        for (int k = 0; k < 27; k++) acc[k] = 0.;
        if (act) {
            double body = 0.;
            rg_gather_slot_general<0>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<1>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<2>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<3>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<4>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<5>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<6>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<7>(m, sK, p.inst_cap, acc, body);
            rg_rhs(p, m, acc, body);
        }
        rg_write_rows(p, m, acc, act, st, lane);
    }
    rg_bulk_wait_read();
}
#endif  // __CUDACC__
