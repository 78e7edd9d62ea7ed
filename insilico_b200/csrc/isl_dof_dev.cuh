// =============================================================================
// isl_dof_dev.cuh -- DoF-object ids per element ON THE DEVICE (included by isl_engine.cu).
//
// Reference: base::dof::generate<FEBasis>(mesh, field) (base/dof/generate.hpp:46-115) through
// base::dof::IndexMap (IndexMap.hpp:221-319) and generateDoFIndicesFromFaces (generateDoFIndicesFromFaces.hpp:169-298):
// for every n-face type in turn (vertices, edges, faces, cell interior) the elements are visited in order, every n-face
// of an element is looked up in a std::map keyed by its sorted vertex tuple, and a face met for the first time receives
// the next `stride` ids.  The host restatement (isl_dof.hpp: dof_generate) does the same with a hash table.
//
// Here, per n-face type: items i = element * nfaces + local face; key = sorted vertex tuple (a vertex, an edge, or the
// vertices of a face: 128 bits, sorted as two 64-bit halves, least significant first); a STABLE radix
// sort of (key, i) puts the first visitor of every face at the head of its run; "first visitor" flags, an exclusive
// scan in visiting order = the reference's running counter.  Same ids as the host version, bit for bit.
// =============================================================================
#pragma once

struct DgTopo { int nfaces, nv; int vert[12][4]; };   // local vertex numbers of every n-face of the current type

__global__ void k_dg_keys(const int32_t* conn, int64_t n_elems, int npe, DgTopo T, uint64_t* khi, uint64_t* klo, uint32_t* idx) {
    const int64_t n = n_elems * T.nfaces;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / T.nfaces; const int f = (int)(i % T.nfaces);
        uint32_t v[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        for (int j = 0; j < T.nv; j++) v[j] = (uint32_t)conn[e * npe + T.vert[f][j]];
        // sorting network for up to four entries (unused ones are 0xffffffff and stay last)
        auto cswap = [](uint32_t& a, uint32_t& b) { if (a > b) { const uint32_t t = a; a = b; b = t; } };
        cswap(v[0], v[1]); cswap(v[2], v[3]); cswap(v[0], v[2]); cswap(v[1], v[3]); cswap(v[1], v[2]);
        khi[i] = ((uint64_t)v[0] << 32) | (uint64_t)(T.nv > 1 ? v[1] : 0u);
        klo[i] = (T.nv > 2) ? (((uint64_t)v[2] << 32) | (uint64_t)v[3]) : 0ull;
        idx[i] = (uint32_t)i;
    }
}
__global__ void k_dg_gather64(const uint64_t* src, const uint32_t* idx, int64_t n, uint64_t* dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}
// head of a run of equal keys -> its own position, else 0 (an inclusive max-scan then gives every member its run start)
__global__ void k_dg_heads(const uint64_t* khi, const uint64_t* klo_orig, const uint32_t* idx, int64_t n, uint32_t* start) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const bool head = (j == 0) || khi[j] != khi[j - 1] || klo_orig[idx[j]] != klo_orig[idx[j - 1]];
        start[j] = head ? (uint32_t)j : 0u;
    }
}
// first visitor of the face of every item, and the "met for the first time" flag in visiting order
__global__ void k_dg_first(const uint32_t* idx, const uint32_t* start, int64_t n, uint32_t* first, uint32_t* fresh) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t i = idx[j], f = idx[start[j]];
        first[i] = f;
        fresh[i] = (f == i) ? 1u : 0u;
    }
}
__global__ void k_dg_write(const uint32_t* first, const uint32_t* rank, int64_t n_elems, int nfaces, int stride, int total, int begin,
                           int64_t next0, int32_t* elem_dof) {
    const int64_t n = n_elems * nfaces;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / nfaces; const int f = (int)(i % nfaces);
        const int64_t base = next0 + (int64_t)rank[first[i]] * stride;   // ids of a shared face: the first owner's, in its local order
        for (int d = 0; d < stride; d++) elem_dof[e * total + begin + f * stride + d] = (int32_t)(base + d);
    }
}
__global__ void k_dg_interior(int64_t n_elems, int per_elem, int total, int begin, int64_t next0, int32_t* elem_dof) {
    const int64_t n = n_elems * per_elem;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        elem_dof[(i / per_elem) * total + begin + (i % per_elem)] = (int32_t)(next0 + i);
}
__global__ void k_dg_copy_conn(const int32_t* conn, int64_t n, int32_t* elem_dof, int* maxid) {
    int m = -1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) { elem_dof[i] = conn[i]; m = max(m, conn[i]); }
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(maxid, m);
}
struct DgMax { __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };
