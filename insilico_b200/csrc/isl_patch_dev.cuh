// =============================================================================
// isl_patch_dev.cuh -- patch formation for the Q1 row kernel ON THE DEVICE (included by isl_engine.cu).
//
// Replaces the host preprocessing of round 1 (recursive coordinate bisection + per-patch element / node lists on the
// host threads: 4-5 s at 256^3 and five full-array copies to the host) for the row-gather kernel.  What the reference
// does at this point is solver.registerFields<FTB>(binder) (base/solver/TripletContainer.hpp:158-301): a one-off per
// numbering.  Everything here is cub sorts / scans and small kernels:
//
//   1. row -> node (the field is isoparametric: DoF = node), lattice coordinate of every row node = position divided by
//      the mean element extent per axis, rounded: robust against node perturbations below half a mesh width, where a
//      coordinate bisection splits a perturbed lattice plane at random.
//   2. box of every row: (Bs x Bo x Bo) lattice cells, Bs along the axis in which consecutive equation numbers advance
//      (long runs of consecutive rows = long contiguous pieces of the value array per patch); patch = box, rows sorted by
//      (box, row) with one stable radix sort.
//   3. element instances of a patch: every (box of row(e, a), e) pair, sorted and made unique (64-bit keys).
//   4. nodes of a patch: every (patch, node) pair of its instances, sorted and made unique; local node index of an
//      instance corner = rank inside the patch's node segment (binary search).
//   5. slot table: the instance in which a row is local node a (atomicCAS: a second claimant means the mesh is not
//      lattice-like -> the caller falls back).
// =============================================================================
#pragma once

// (cub/device/device_run_length_encode.cuh and device_scan.cuh are included at the top of isl_engine.cu)

__global__ void k_pd_row_node(const int32_t* node_eqn, int64_t n_nodes, int32_t* row_node) {
    for (int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; nd < n_nodes; nd += (int64_t)gridDim.x * blockDim.x)
        if (node_eqn[nd] >= 0) row_node[node_eqn[nd]] = (int32_t)nd;
}
// mean element extent per axis from a sample of the elements: sums[0..2] += max - min, sums[3] += 1
__global__ void k_pd_extent(const double* coords, const int32_t* conn, int64_t n, int64_t stride, double* sums) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k * stride < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = k * stride;
        for (int d = 0; d < 3; d++) {
            double mn = 1e300, mx = -1e300;
            for (int a = 0; a < 8; a++) { const double v = coords[(size_t)conn[e * 8 + a] * 3 + d]; mn = fmin(mn, v); mx = fmax(mx, v); }
            atomicAdd(sums + d, mx - mn);
        }
        atomicAdd(sums + 3, 1.0);
    }
}
// along which axis do consecutive equation numbers advance?  votes[axis]
__global__ void k_pd_axis_votes(const double* coords, const int32_t* row_node, int64_t n_rows, int64_t stride, const double* hmean,
                                unsigned long long* votes) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; (k + 1) * stride < n_rows; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = k * stride;
        const int32_t a = row_node[r], b = row_node[r + 1];
        if (a < 0 || b < 0) continue;
        int best = -1, nz = 0;
        for (int d = 0; d < 3; d++)
            if (fabs(coords[(size_t)b * 3 + d] - coords[(size_t)a * 3 + d]) > 0.25 * hmean[d]) { nz++; best = d; }
        if (nz == 1) atomicAdd(votes + best, 1ull);
    }
}
struct PdGrid { double lo[3], h[3]; int B[3], NB[3]; };
__global__ void k_pd_row_keys(const double* coords, const int32_t* row_node, int64_t n_rows, PdGrid g, int32_t* key, int32_t* idx) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int32_t nd = row_node[r];
        int b[3] = {0, 0, 0};
        if (nd >= 0)
            for (int d = 0; d < 3; d++) {
                const long long cell = (long long)floor((coords[(size_t)nd * 3 + d] - g.lo[d]) / g.h[d] + 0.5);
                b[d] = (int)min((long long)g.NB[d] - 1, max(0ll, cell / g.B[d]));
            }
        key[r] = (b[2] * g.NB[1] + b[1]) * g.NB[0] + b[0];
        idx[r] = (int32_t)r;
    }
}
// patch index of every row from the run offsets of the sorted keys
__global__ void k_pd_patch_of_row(const int32_t* p_row_off, int n_patches, const int32_t* rows, int32_t* patch_of_row, int32_t* rank_of_row) {
    const int p = blockIdx.x;
    if (p >= n_patches) return;
    for (int r = p_row_off[p] + threadIdx.x; r < p_row_off[p + 1]; r += blockDim.x) {
        patch_of_row[rows[r]] = p;
        rank_of_row[rows[r]] = r - p_row_off[p];
    }
}
__global__ void k_pd_inst_keys(const int32_t* elem_eqn, int64_t n_elems, const int32_t* patch_of_row, uint64_t* keys) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_elems * 8; t += (int64_t)gridDim.x * blockDim.x) {
        const int32_t q = elem_eqn[t];
        keys[t] = (q >= 0) ? (((uint64_t)(uint32_t)patch_of_row[q] << 32) | (uint64_t)(t >> 3)) : ~0ull;
    }
}
// first index whose key has a high word >= p, for p = 0..n_patches
__global__ void k_pd_segment_offsets(const uint64_t* keys, int64_t n, int n_patches, int32_t* off) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n_patches) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((keys[mid] >> 32) < (uint64_t)p) lo = mid + 1; else hi = mid; }
    off[p] = (int32_t)lo;
}
__global__ void k_pd_node_keys(const uint64_t* inst_keys, int64_t n_inst, const int32_t* conn, uint64_t* keys, int32_t* inst_elem) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_inst; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = inst_keys[i];
        const int64_t e = (int64_t)(k & 0xffffffffu);
        inst_elem[i] = (int32_t)e;
        for (int a = 0; a < 8; a++) keys[i * 8 + a] = (k & 0xffffffff00000000ull) | (uint64_t)(uint32_t)conn[e * 8 + a];
    }
}
__global__ void k_pd_nodes_from_keys(const uint64_t* keys, int64_t n, int32_t* nodes) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) nodes[i] = (int32_t)(keys[i] & 0xffffffffu);
}
// local node indices and slot table of every instance; err[0]: not lattice-like, err[1]: a patch has too many nodes
__global__ void k_pd_instances(const uint64_t* inst_keys, int64_t n_inst, const int32_t* conn, const int32_t* elem_eqn,
                               const int32_t* p_inst_off, const int32_t* p_node_off, const int32_t* p_row_off, const int32_t* nodes,
                               const int32_t* patch_of_row, const int32_t* rank_of_row, uint16_t* lnode, uint16_t* rslot, int* err) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_inst; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = inst_keys[i];
        const int p = (int)(k >> 32);
        const int64_t e = (int64_t)(k & 0xffffffffu);
        const int li = (int)(i - p_inst_off[p]);
        const int n0 = p_node_off[p], n1 = p_node_off[p + 1];
        if (n1 - n0 >= 65535 || li >= 65535) { err[1] = 1; continue; }
        int32_t q[8];
        for (int a = 0; a < 8; a++) q[a] = elem_eqn[e * 8 + a];
        for (int a = 0; a < 8; a++) {
            const int32_t nd = conn[e * 8 + a];
            int lo = n0, hi = n1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (nodes[mid] < nd) lo = mid + 1; else hi = mid; }
            lnode[i * 8 + a] = (uint16_t)(lo - n0);
            if (q[a] < 0) continue;
            for (int b = 0; b < a; b++) if (q[b] == q[a]) err[0] = 1;   // tied numbering inside an element
            if (patch_of_row[q[a]] != p) continue;
            unsigned short* s = reinterpret_cast<unsigned short*>(rslot) + ((size_t)(p_row_off[p] + rank_of_row[q[a]]) * 8 + a);
            if (atomicCAS(s, (unsigned short)0xffff, (unsigned short)li) != 0xffff) err[0] = 1;   // two elements see the row as local node a
        }
    }
}
