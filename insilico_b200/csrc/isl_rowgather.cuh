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
};
static_assert(sizeof(RowMeta) == 48, "RowMeta is read as three 16-byte words");

// tables: device = the engine's constant memory, host emulation = plain arrays filled by the test harness
#ifdef __CUDA_ARCH__
#define RG_AFF(i) c_q1_aff[i]
#define RG_NSUM(i) c_q1_Nsum[i]
#else
static double rg_host_aff[6 * 36];
static double rg_host_nsum[8];
#define RG_AFF(i) rg_host_aff[i]
#define RG_NSUM(i) rg_host_nsum[i]
#endif

// phase 1: the six numbers D_c and det J w of an affine element from its hierarchic nodes 0, 1 (xi), 3 (eta), 4 (zeta)
RG_HD void rg_instance(const double* x0, const double* x1, const double* x3, const double* x4, double factor, double w,
                       double (&D)[6], double& dw) {
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
    const double s = (factor * w) / det;
    dw = det * w;
    D[0] = s * (co[0][0] * co[0][0] + co[1][0] * co[1][0] + co[2][0] * co[2][0]);
    D[1] = s * (co[0][1] * co[0][1] + co[1][1] * co[1][1] + co[2][1] * co[2][1]);
    D[2] = s * (co[0][2] * co[0][2] + co[1][2] * co[1][2] + co[2][2] * co[2][2]);
    D[3] = s * (co[0][0] * co[0][1] + co[1][0] * co[1][1] + co[2][0] * co[2][1]);
    D[4] = s * (co[0][0] * co[0][2] + co[1][0] * co[1][2] + co[2][0] * co[2][2]);
    D[5] = s * (co[0][1] * co[0][2] + co[1][1] * co[1][2] + co[2][1] * co[2][2]);
}

// phase 2: the eight entries that the element with the row node as local node A contributes (A compile-time)
template <int A>
RG_HD void rg_add_slot(const double (&D)[6], double dw, double (&acc)[27], double& body) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int b = 0; b < 8; b++) {
        double v = D[0] * RG_AFF(sym_idx(A, b));
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int c = 1; c < 6; c++) v = fma(D[c], RG_AFF(c * 36 + sym_idx(A, b)), v);
        acc[rg_kidx(A, b)] += v;
    }
    body = fma(dw, RG_NSUM(A), body);
}

// Signed-sum form of the same entries (SS variant).  On the reference cube the trilinear integrals are exact rationals:
//   int d_x N_a d_x N_b = s_x(a) s_x(b) m_y m_z,  m = 1/3 (a, b on the same side of that axis) or 1/6,  s = -1 / +1,
//   int (d_x N_a d_y N_b + d_y N_a d_x N_b) = +- (1/2) m_z s_x(a) s_y(a)  when a, b agree on both or differ on both
//                                              of x, y (+ / -), else 0
// (the 2x2x2 Gauss rule integrates them exactly, so this equals the quadrature of the reference up to rounding).
// Phase 1 stores, per element, E_c {1/9, 1/18, 1/36} (c = xx, yy, zz) and E_c {1/6, 1/12} (c = xy, xz, yz) with
// E_c = (kappa / det J) (cof^T cof)_c: 15 numbers; every matrix entry is then a signed sum of 3 to 6 of them with
// compile-time signs and selections: 288 additions per row instead of 384 fused multiply-adds + 64 additions, and no
// constant-table loads.
RG_HD void rg_instance_ss(const double* x0, const double* x1, const double* x3, const double* x4, double factor, double w,
                          double (&T)[15], double& dw) {
    double D[6];
    rg_instance(x0, x1, x3, x4, factor, w, D, dw);   // D_c = (factor w / det) (cof^T cof)_c with w = 1/8
    const double inv_w = 1.0 / w;
    for (int c = 0; c < 3; c++) {
        const double e = D[c] * inv_w;
        T[c * 3] = e * (1.0 / 9.0); T[c * 3 + 1] = e * (1.0 / 18.0); T[c * 3 + 2] = e * (1.0 / 36.0);
    }
    for (int c = 3; c < 6; c++) {
        const double e = D[c] * inv_w;
        T[9 + (c - 3) * 2] = e * (1.0 / 6.0); T[9 + (c - 3) * 2 + 1] = e * (1.0 / 12.0);
    }
}
RG_HD constexpr int rg_bit(int axis, int a) { return axis == 0 ? rg_bx(a) : axis == 1 ? rg_by(a) : rg_bz(a); }
RG_HD constexpr bool rg_same(int axis, int a, int b) { return rg_bit(axis, a) == rg_bit(axis, b); }
RG_HD constexpr int rg_sgn(int axis, int a) { return rg_bit(axis, a) ? 1 : -1; }
template <int A>
RG_HD void rg_add_slot_ss(const double (&T)[15], double dw, double (&acc)[27], double& body) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int b = 0; b < 8; b++) {
        double v = acc[rg_kidx(A, b)];
        // diagonal terms c = 0 (xx), 1 (yy), 2 (zz): sign s_c(a) s_c(b), magnitude by the two other axes
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int c = 0; c < 3; c++) {
            const int o1 = (c + 1) % 3, o2 = (c + 2) % 3;
            const int mi = (rg_same(o1, A, b) ? 0 : 1) + (rg_same(o2, A, b) ? 0 : 1);
            const double t = T[c * 3 + mi];
            v = (rg_sgn(c, A) * rg_sgn(c, b) > 0) ? v + t : v - t;
        }
        // cross terms c = 3 (xy), 4 (xz), 5 (yz)
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int c = 3; c < 6; c++) {
            const int p = (c == 5) ? 1 : 0, q = (c == 3) ? 1 : 2, r = 3 - p - q;
            const bool ep = rg_same(p, A, b), eq = rg_same(q, A, b);
            if (ep == eq) {
                const double t = T[9 + (c - 3) * 2 + (rg_same(r, A, b) ? 0 : 1)];
                v = ((rg_sgn(p, A) * rg_sgn(q, A) > 0) == ep) ? v + t : v - t;
            }
        }
        acc[rg_kidx(A, b)] = v;
    }
    body = fma(dw, RG_NSUM(A), body);
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
// positions, whether a neighbour is CONSTRAINED.  inst_elem points at the patch's instance list.  Returns false when
// the row is not eligible (see header).  nbn[k] = global node of stencil neighbour k or -1.
RG_HD bool rg_row_tables(int32_t g, const uint16_t* slot, const int32_t* inst_elem, const int32_t* conn,
                         const int32_t* node_eqn, const uint8_t* status, const int64_t* rowptr, const int32_t* col,
                         RowMeta& m, int32_t (&nbn)[27], bool& constrained_nb) {
    for (int k = 0; k < 27; k++) { nbn[k] = -1; m.pos[k] = 0xff; }
    for (int a = 0; a < 8; a++) {
        m.slot[a] = slot[a];
        if (slot[a] == 0xffff) continue;
        const int32_t* ce = conn + (size_t)inst_elem[slot[a]] * 8;
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
    return true;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device side
struct RowsParams {
    const double* coords;
    const int32_t* p_inst_off; const int32_t* p_row_off; const int32_t* p_node_off;
    const int32_t* rows;       // global row id of every owned local row
    const int32_t* nodes;      // global node id of every local node
    const uint16_t* i_lnode;   // [inst][8] local node index
    const RowMeta* meta;       // per owned row (same index as rows)
    const int64_t* rowstart;   // per owned row: rowptr[row]
    const int32_t* lift_nodes; // [flagged rows][27] neighbour node ids
    const uint8_t* status; const double* presc; const double* values;
    double* val; double* rhs;
    double factor; int incremental; int store_mode;
    int body; double f0;
    int node_cap, inst_cap;
};

// preprocessing: one CTA per patch, threads over its rows.  pass 0 fills meta / rowstart and numbers the rows that
// have CONSTRAINED neighbours (counter[0]); pass 1 writes their neighbour node ids.  err[0] != 0: some row is not eligible.
__global__ void k_row_meta(int pass, const int32_t* p_row_off, const int32_t* p_inst_off, const int32_t* rows,
                           const uint16_t* rslot, const int32_t* inst_elem, const int32_t* conn, const int32_t* node_eqn,
                           const uint8_t* status, const int64_t* rowptr, const int32_t* col, RowMeta* meta,
                           int64_t* rowstart, int32_t* lift_nodes, int* counter, int* err) {
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
            rowstart[r0 + r] = rowptr[g];
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


// matrix write-out of a warp's 32 rows: sixteen rows at a time through the warp's staging buffer into CSR order, then
// stores along the rows.  When the sixteen rows are consecutive in the CSR arrays (the usual case: consecutive
// equation numbers) the staged block is one contiguous piece of the value array and is streamed with all 32 lanes.
__device__ __forceinline__ void rg_write_rows(const RowsParams& p, const RowMeta& m, const double (&acc)[27], bool act, int myn,
                                              int64_t rs, double* st, int lane) {
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
        const bool mine = (lane >> 4) == h;
        int incl = mine ? myn : 0;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d, 16);
            if ((lane & 15) >= d) incl += v;
        }
        const int off = incl - (mine ? myn : 0);
        if (mine && act) {
            if (m.nnz & 0x80) {
#pragma unroll
                for (int k = 0; k < 27; k++)
                    if (k < myn) st[off + k] = 0.;
            }
#pragma unroll
            for (int k = 0; k < 27; k++)
                if (m.pos[k] != 0xff) st[off + m.pos[k]] = acc[k];
        }
        const int64_t rs_next = __shfl_down_sync(0xffffffffu, rs, 1);
        const int n_next = __shfl_down_sync(0xffffffffu, myn, 1);
        const bool ok = !mine || (lane & 15) == 15 || n_next == 0 || rs_next == rs + myn;
        const bool contiguous = __all_sync(0xffffffffu, ok);
        __syncwarp();
        if (contiguous) {
            const int total = __shfl_sync(0xffffffffu, incl, h * 16 + 15);
            const int64_t rs0 = __shfl_sync(0xffffffffu, rs, h * 16);
            if (p.store_mode) {
                for (int q = lane; q < total; q += 32) p.val[rs0 + q] = st[q];
            } else {
                for (int q = lane; q < total; q += 32) p.val[rs0 + q] += st[q];
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < 16; j++) {
                const int src = h * 16 + j;
                const int64_t rsj = __shfl_sync(0xffffffffu, rs, src);
                const int nj = __shfl_sync(0xffffffffu, myn, src);
                const int oj = __shfl_sync(0xffffffffu, off, src);
                if (lane < nj) {
                    if (p.store_mode) p.val[rsj + lane] = st[oj + lane];
                    else p.val[rsj + lane] += st[oj + lane];
                }
            }
        }
        __syncwarp();
    }
}

template <int A>
__device__ __forceinline__ void rg_gather_slot(const RowMeta& m, const double* sD, int cap, double (&acc)[27], double& body) {
    const int s = m.slot[A];
    if (s != 0xffff) {
        double D[6];
#pragma unroll
        for (int c = 0; c < 6; c++) D[c] = sD[c * cap + s];
        rg_add_slot<A>(D, sD[6 * cap + s], acc, body);
    }
}

template <int A>
__device__ __forceinline__ void rg_gather_slot_ss(const RowMeta& m, const double* sD, int cap, double (&acc)[27], double& body) {
    const int s = m.slot[A];
    if (s != 0xffff) {
        double T[15];
#pragma unroll
        for (int c = 0; c < 15; c++) T[c] = sD[c * cap + s];
        rg_add_slot_ss<A>(T, sD[15 * cap + s], acc, body);
    }
}

// one CTA per patch.  Shared memory: sD[7 (SS: 16)][inst_cap] | (sX[node_cap][3]  aliased after phase 1 by  stage[NT/32][16*27])
template <int NT, int MINB, bool SS = false>
__global__ void __launch_bounds__(NT, MINB) k_q1hex_rows_affine(const RowsParams p) {
    extern __shared__ double smem[];
    double* sD = smem;
    double* sX = sD + (size_t)(SS ? 16 : 7) * p.inst_cap;
    double* stage = sX;
    const int tid = threadIdx.x;
    const int pid = blockIdx.x;
    const int r0 = p.p_row_off[pid], nrows = p.p_row_off[pid + 1] - r0;
    const int n0 = p.p_node_off[pid], nnodes = p.p_node_off[pid + 1] - n0;
    const int e0 = p.p_inst_off[pid], ninst = p.p_inst_off[pid + 1] - e0;
    {
        // the patch's element instances and row tables are contiguous: pull them into L2 while the coordinates load
        const char* b1 = reinterpret_cast<const char*>(p.i_lnode + (size_t)e0 * 8);
        for (size_t o = (size_t)tid * 128; o < (size_t)ninst * 16; o += (size_t)NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b1 + o));
        const char* b2 = reinterpret_cast<const char*>(p.meta + r0);
        for (size_t o = (size_t)tid * 128; o < (size_t)nrows * sizeof(RowMeta); o += (size_t)NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + o));
    }
    // nodal coordinates of the patch -> shared memory
    {
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
    __syncthreads();
    // phase 1: element instances
    const double w = c_q1_w[0];
    for (int i = tid; i < ninst; i += NT) {
        const int4 l4 = __ldg(reinterpret_cast<const int4*>(p.i_lnode + (size_t)(e0 + i) * 8));
        const int n0l = l4.x & 0xffff, n1l = (unsigned)l4.x >> 16, n3l = (unsigned)l4.y >> 16, n4l = l4.z & 0xffff;
        if (SS) {
            double T[15], dw;
            rg_instance_ss(sX + n0l * 3, sX + n1l * 3, sX + n3l * 3, sX + n4l * 3, p.factor, w, T, dw);
#pragma unroll
            for (int c = 0; c < 15; c++) sD[c * p.inst_cap + i] = T[c];
            sD[15 * p.inst_cap + i] = dw;
        } else {
            double D[6], dw;
            rg_instance(sX + n0l * 3, sX + n1l * 3, sX + n3l * 3, sX + n4l * 3, p.factor, w, D, dw);
#pragma unroll
            for (int c = 0; c < 6; c++) sD[c * p.inst_cap + i] = D[c];
            sD[6 * p.inst_cap + i] = dw;
        }
    }
    __syncthreads();  // sD complete; sX is dead from here on (stage aliases it)
    // phase 2: owned rows
    const int lane = tid & 31, warp = tid >> 5;
    double* st = stage + (size_t)warp * (16 * 27);
    for (int rb = 0; rb < nrows; rb += NT) {
        const int r = rb + tid;
        const bool act = r < nrows;
        RowMeta m;
        int64_t rs = 0;
        double acc[27];
#pragma unroll
        for (int k = 0; k < 27; k++) acc[k] = 0.;
        int myn = 0;
        if (act) {
            const int4* mp = reinterpret_cast<const int4*>(p.meta + r0 + r);
            int4* md = reinterpret_cast<int4*>(&m);
            md[0] = __ldg(mp); md[1] = __ldg(mp + 1); md[2] = __ldg(mp + 2);
            rs = __ldg(p.rowstart + r0 + r);
            myn = m.nnz & 0x7f;
            double body = 0.;
            if (SS) {
                rg_gather_slot_ss<0>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot_ss<1>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot_ss<2>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot_ss<3>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot_ss<4>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot_ss<5>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot_ss<6>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot_ss<7>(m, sD, p.inst_cap, acc, body);
            } else {
                rg_gather_slot<0>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot<1>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot<2>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot<3>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot<4>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot<5>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot<6>(m, sD, p.inst_cap, acc, body);
                rg_gather_slot<7>(m, sD, p.inst_cap, acc, body);
            }
            // right-hand side: Dirichlet lift of CONSTRAINED stencil neighbours (assembleMatrix.hpp:56-130), body force
            double lift = 0.;
            if (m.lift >= 0) {
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
            if (v != 0.) { const int32_t g = __ldg(p.rows + r0 + r); p.rhs[g] += v; }
        }
        // matrix: sixteen rows at a time through the warp's staging buffer into CSR order, then row-contiguous stores
        rg_write_rows(p, m, acc, act, myn, rs, st, lane);
    }
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
    extern __shared__ double smem[];
    double* sK = smem;
    double* sX = sK + (size_t)44 * p.inst_cap;
    double* stage = sX;
    const int tid = threadIdx.x;
    const int pid = blockIdx.x;
    const int r0 = p.p_row_off[pid], nrows = p.p_row_off[pid + 1] - r0;
    const int n0 = p.p_node_off[pid], nnodes = p.p_node_off[pid + 1] - n0;
    const int e0 = p.p_inst_off[pid], ninst = p.p_inst_off[pid + 1] - e0;
    {
        const char* b1 = reinterpret_cast<const char*>(p.i_lnode + (size_t)e0 * 8);
        for (size_t o = (size_t)tid * 128; o < (size_t)ninst * 16; o += (size_t)NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b1 + o));
        const char* b2 = reinterpret_cast<const char*>(p.meta + r0);
        for (size_t o = (size_t)tid * 128; o < (size_t)nrows * sizeof(RowMeta); o += (size_t)NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + o));
    }
    for (int n = tid; n < nnodes; n += NT) {
        const double* c = p.coords + (size_t)__ldg(p.nodes + n0 + n) * 3;
        sX[n * 3] = __ldg(c); sX[n * 3 + 1] = __ldg(c + 1); sX[n * 3 + 2] = __ldg(c + 2);
    }
    __syncthreads();
    // phase 1: local matrices of the element instances
    for (int i = tid; i < ninst; i += NT) {
        const int4 l4 = __ldg(reinterpret_cast<const int4*>(p.i_lnode + (size_t)(e0 + i) * 8));
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
    }
    __syncthreads();  // sK complete; sX is dead from here on (stage aliases it)
    // phase 2: owned rows (identical to k_q1hex_rows_affine apart from the slot routine)
    const int lane = tid & 31, warp = tid >> 5;
    double* st = stage + (size_t)warp * (16 * 27);
    for (int rb = 0; rb < nrows; rb += NT) {
        const int r = rb + tid;
        const bool act = r < nrows;
        RowMeta m;
        int64_t rs = 0;
        double acc[27];
#pragma unroll
        for (int k = 0; k < 27; k++) acc[k] = 0.;
        int myn = 0;
        if (act) {
            const int4* mp = reinterpret_cast<const int4*>(p.meta + r0 + r);
            int4* md = reinterpret_cast<int4*>(&m);
            md[0] = __ldg(mp); md[1] = __ldg(mp + 1); md[2] = __ldg(mp + 2);
            rs = __ldg(p.rowstart + r0 + r);
            myn = m.nnz & 0x7f;
            double body = 0.;
            rg_gather_slot_general<0>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<1>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<2>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<3>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<4>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<5>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<6>(m, sK, p.inst_cap, acc, body);
            rg_gather_slot_general<7>(m, sK, p.inst_cap, acc, body);
            double lift = 0.;
            if (m.lift >= 0) {
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
            if (v != 0.) { const int32_t g = __ldg(p.rows + r0 + r); p.rhs[g] += v; }
        }
        rg_write_rows(p, m, acc, act, myn, rs, st, lane);
    }
}
#endif  // __CUDACC__
