// Host-side DoF handling of the B200 assembly engine on flat arrays.
//
// Produces the same DoF-object ids, boundary list and equation numbers as the reference
//   base/dof/IndexMap.hpp:221-280, base/dof/copyConnectivity.hpp:34-62,
//   base/dof/generateDoFIndicesFromFaces.hpp:169-298, base/mesh/createBoundaryFromUnstructured.hpp:55-106,
//   base/dof/constrainBoundary.hpp:49-123, base/fe/Policies.hpp:99-204, base/dof/numbering.hpp:44-68
// but with a different mechanism: an open-addressing hash table keyed by the sorted vertex tuple of an
// n-face replaces the reference's std::map, so one pass over the elements is O(#faces).
#pragma once
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <vector>

#include "isl_tables.hpp"

namespace isl {

struct FaceKey {
    int32_t v[4];
    bool operator==(const FaceKey& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
    bool operator<(const FaceKey& o) const { return std::lexicographical_compare(v, v + 4, o.v, o.v + 4); }
};

// first-appearance table: key -> payload (element, face number), linear probing, power-of-two capacity
class FaceTable {
public:
    explicit FaceTable(size_t expected) {
        size_t cap = 16;
        while (cap < expected * 2) cap <<= 1;
        mask_ = cap - 1;
        keys_.resize(cap);
        payload_.assign(cap, kEmpty);
    }
    // returns slot; inserted=true if the key was new
    size_t find_or_insert(const FaceKey& k, int64_t payload, bool& inserted) {
        size_t h = hash(k) & mask_;
        for (;;) {
            if (payload_[h] == kEmpty) { keys_[h] = k; payload_[h] = payload; inserted = true; return h; }
            if (keys_[h] == k) { inserted = false; return h; }
            h = (h + 1) & mask_;
        }
    }
    int64_t payload(size_t slot) const { return payload_[slot]; }
    int64_t& payload_ref(size_t slot) { return payload_[slot]; }
    const FaceKey& key(size_t slot) const { return keys_[slot]; }
    size_t capacity() const { return mask_ + 1; }

    static constexpr int64_t kEmpty = INT64_MIN;

private:
    static size_t hash(const FaceKey& k) {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (int i = 0; i < 4; i++) {
            h ^= (uint64_t)(uint32_t)k.v[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
            h *= 0xFF51AFD7ED558CCDull;
            h ^= h >> 33;
        }
        return (size_t)h;
    }
    size_t mask_;
    std::vector<FaceKey> keys_;
    std::vector<int64_t> payload_;
};

inline int nface_num_vertices(int shape, int nf) {
    if (nf == VERTEX) return 1;
    if (nf == EDGE) return 2;
    if (nf == FACE) return topology(shape).face_nv;
    return n_subfaces(shape, VERTEX);
}
inline int nface_vertex(int shape, int nf, int f, int j) {
    const Topology& T = topology(shape);
    if (nf == VERTEX) return f;
    if (nf == EDGE) return T.edge[f][j];
    if (nf == FACE) return T.face[f][j];
    return j;
}
inline FaceKey make_key(int shape, int nf, int f, const int32_t* elem_conn) {
    FaceKey k; k.v[0] = k.v[1] = k.v[2] = k.v[3] = -1;
    const int nv = nface_num_vertices(shape, nf);
    for (int j = 0; j < nv && j < 4; j++) k.v[j] = elem_conn[nface_vertex(shape, nf, f, j)];
    std::sort(k.v, k.v + std::min(nv, 4));
    return k;
}

// DoF-object ids per element.  npe = nodes per geometry element.
inline int64_t dof_generate(int shape, int geom_deg, int64_t n_elems, const int32_t* conn, int fe_deg,
                            int32_t* elem_dof) {
    const Basis geom(shape, geom_deg);
    const int npe = geom.nfun;
    const FELayout L = fe_layout(shape, fe_deg);
    if (fe_deg == geom_deg) {  // isoparametric: DoF id = node id
        int64_t n = 0;
        for (int64_t e = 0; e < n_elems; e++)
            for (int a = 0; a < npe; a++) {
                const int32_t id = conn[e * npe + a];
                elem_dof[e * L.total + a] = id;
                n = std::max<int64_t>(n, (int64_t)id + 1);
            }
        return n;
    }
    const int dim = shape_dim(shape);
    int64_t next = 0;
    for (int nf = 0; nf <= dim; nf++) {
        const int stride = L.per[nf];
        if (stride == 0) continue;
        const int nfaces = L.count[nf];
        if (nf == dim) {  // interior DoFs, private to the element
            for (int64_t e = 0; e < n_elems; e++)
                for (int d = 0; d < stride * nfaces; d++) elem_dof[e * L.total + L.begin[nf] + d] = (int32_t)next++;
            continue;
        }
        FaceTable table((size_t)n_elems * nfaces);
        for (int64_t e = 0; e < n_elems; e++)
            for (int f = 0; f < nfaces; f++) {
                bool fresh;
                const FaceKey key = make_key(shape, nf, f, conn + e * npe);
                const size_t slot = table.find_or_insert(key, next, fresh);
                // ids of a shared n-face are taken over in the first owner's local order (no re-orientation)
                const int64_t base = fresh ? next : table.payload(slot);
                if (fresh) next += stride;
                for (int d = 0; d < stride; d++) elem_dof[e * L.total + L.begin[nf] + f * stride + d] = (int32_t)(base + d);
            }
    }
    return next;
}

// boundary faces = (dim-1)-faces met exactly once, listed in ascending order of their sorted vertex tuple
inline void mesh_boundary(int shape, int geom_deg, int64_t n_elems, const int32_t* conn,
                          std::vector<int64_t>& pairs) {
    const Basis geom(shape, geom_deg);
    const int npe = geom.nfun, surf = shape_dim(shape) - 1;
    const int nfaces = n_subfaces(shape, surf);
    FaceTable table((size_t)n_elems * nfaces);
    // payload = first (element*nfaces+face); a second visit marks the slot interior (-2 - payload)
    for (int64_t e = 0; e < n_elems; e++)
        for (int f = 0; f < nfaces; f++) {
            bool fresh;
            const size_t slot = table.find_or_insert(make_key(shape, surf, f, conn + e * npe), e * nfaces + f, fresh);
            if (!fresh) { int64_t& p = table.payload_ref(slot); if (p >= 0) p = -2 - p; }
        }
    std::vector<std::pair<FaceKey, int64_t>> b;
    for (size_t s = 0; s < table.capacity(); s++) if (table.payload(s) >= 0 && table.payload(s) != FaceTable::kEmpty) b.emplace_back(table.key(s), table.payload(s));
    std::sort(b.begin(), b.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    pairs.clear();
    for (auto& kv : b) { pairs.push_back(kv.second / nfaces); pairs.push_back(kv.second % nfaces); }
}

// local DoF numbers lying on face `face_no` of the surface n-face type (vertices, edge DoFs, face DoFs)
inline void face_local_dofs(int shape, int fe_deg, int nf, int face_no, std::vector<int>& out) {
    const FELayout L = fe_layout(shape, fe_deg);
    const Topology& T = topology(shape);
    out.clear();
    if (nf == VERTEX) { out.push_back(face_no); return; }
    if (nf == EDGE) {
        out.push_back(T.edge[face_no][0]); out.push_back(T.edge[face_no][1]);
        for (int d = 0; d < L.per[EDGE]; d++) out.push_back(L.begin[EDGE] + face_no * L.per[EDGE] + d);
        return;
    }
    if (nf == FACE) {
        for (int v = 0; v < T.face_nv; v++) out.push_back(T.face[face_no][v]);
        for (int e = 0; e < T.face_nv; e++) {
            const int en = T.face_edge[face_no][e], sg = T.face_edge_sign[face_no][e], es = L.per[EDGE];
            for (int d = 0; d < es; d++) out.push_back(L.begin[EDGE] + en * es + (sg > 0 ? d : es - 1 - d));
        }
        for (int d = 0; d < L.per[FACE]; d++) out.push_back(L.begin[FACE] + face_no * L.per[FACE] + d);
        return;
    }
    for (int i = 0; i < L.total; i++) out.push_back(i);
}

// DoF objects visited by constrainBoundary and the physical position of their support points
inline void boundary_dofs(int shape, int geom_deg, int dim, const double* coords, const int32_t* conn, int fe_deg,
                          const int32_t* elem_dof, int64_t n_pairs, const int64_t* pairs, std::vector<int32_t>& obj,
                          std::vector<double>& x) {
    const Basis geom(shape, geom_deg), fe(shape, fe_deg);
    const FELayout L = fe_layout(shape, fe_deg);
    std::vector<double> sp((size_t)fe.nfun * fe.dim), N(geom.nfun);
    fe.support(sp.data());
    const int surf = shape_dim(shape) - 1;
    std::vector<int> loc;
    obj.clear(); x.clear();
    for (int64_t b = 0; b < n_pairs; b++) {
        const int64_t e = pairs[2 * b];
        face_local_dofs(shape, fe_deg, surf, (int)pairs[2 * b + 1], loc);
        for (int l : loc) {
            geom.eval(&sp[(size_t)l * fe.dim], N.data(), nullptr);
            double xx[3] = {0., 0., 0.};
            for (int a = 0; a < geom.nfun; a++) {
                const int32_t node = conn[e * geom.nfun + a];
                for (int d = 0; d < dim; d++) xx[d] += coords[(size_t)node * dim + d] * N[a];
            }
            obj.push_back(elem_dof[e * L.total + l]);
            for (int d = 0; d < dim; d++) x.push_back(xx[d]);
        }
    }
}

// ---- surface elements of boundary faces ------------------------------------------------------------------------------
// shape of the faces of a shape (base/shape.hpp FaceShape)
inline int face_shape(int shape) { return shape == HEX ? QUAD : (shape == TET ? TRI : LINE); }

// base::mesh::generateBoundaryMesh (base/mesh/generateBoundaryMesh.hpp:279-432, without the optional triangulation): one
// surface element per (element, face number) pair.  Its nodes are the geometry nodes of the domain element that
// FaceExtraction lists for the face, in that order (= the order the Lagrange shape functions of the face shape expect);
// with every node it keeps the node's coordinate in the parameter space of the domain element (base/mesh/SurfaceElement.hpp).
// domain_elem[n_pairs], surf_x[n_pairs * P * dim], surf_param[n_pairs * P * dim]; returns P.
inline int boundary_surface(int shape, int geom_deg, int dim, const double* coords, const int32_t* conn, int64_t n_pairs,
                            const int64_t* pairs, int32_t* domain_elem, double* surf_x, double* surf_param) {
    const Basis geom(shape, geom_deg);
    std::vector<double> sp((size_t)geom.nfun * geom.dim);
    geom.support(sp.data());
    const int surf = shape_dim(shape) - 1;
    std::vector<int> loc;
    face_local_dofs(shape, geom_deg, surf, 0, loc);
    const int P = (int)loc.size();
    if (!domain_elem) return P;
    for (int64_t b = 0; b < n_pairs; b++) {
        const int64_t e = pairs[2 * b];
        face_local_dofs(shape, geom_deg, surf, (int)pairs[2 * b + 1], loc);
        domain_elem[b] = (int32_t)e;
        for (int p = 0; p < P; p++) {
            const int32_t node = conn[e * geom.nfun + loc[p]];
            for (int d = 0; d < dim; d++) {
                surf_x[((size_t)b * P + p) * dim + d] = coords[(size_t)node * dim + d];
                surf_param[((size_t)b * P + p) * dim + d] = sp[(size_t)loc[p] * dim + d];
            }
        }
    }
    return P;
}

// un-normalised normal of a surface element at a point: cross product of the columns of its Jacobi matrix
// (base/geometry.hpp:256-273, base/linearAlgebra.hpp:75-99); J[d][a] = sum_p x_p[d] dN_p/deta_a
inline void surface_normal_raw(int dim, int P, const double* xs, const double* dN, double* nrm) {
    double J[3][2] = {{0., 0.}, {0., 0.}, {0., 0.}};
    const int ld = dim - 1;
    for (int p = 0; p < P; p++)
        for (int d = 0; d < dim; d++)
            for (int a = 0; a < ld; a++) J[d][a] += xs[p * dim + d] * dN[p * ld + a];
    if (dim == 3) {
        nrm[0] = J[1][0] * J[2][1] - J[2][0] * J[1][1];
        nrm[1] = J[2][0] * J[0][1] - J[0][0] * J[2][1];
        nrm[2] = J[0][0] * J[1][1] - J[1][0] * J[0][1];
    } else {
        nrm[0] = J[1][0]; nrm[1] = -J[0][0];
    }
}

// physical position, unit normal and surface metric at the points of SurfaceQuadrature<quad_deg> of every surface element
// (what base::asmb::NeumannForce hands to the caller's force function, base/asmb/NeumannForce.hpp:152-163):
// x[n_surf * nq * dim], normal[n_surf * nq * dim], detg[n_surf * nq] (any may be null); returns nq
inline int surface_points(int surf_shape, int geom_deg, int dim, int64_t n_surf, const double* surf_x, int quad_deg,
                          double* x, double* normal, double* detg) {
    const Basis sg(surf_shape, geom_deg);
    const Rule R = make_rule(surf_shape, quad_deg);
    const int P = sg.nfun, ld = dim - 1;
    std::vector<double> N((size_t)R.n * P), dN((size_t)R.n * P * ld);
    for (int q = 0; q < R.n; q++) sg.eval(&R.p[(size_t)q * ld], &N[(size_t)q * P], &dN[(size_t)q * P * ld]);
    if (!surf_x) return R.n;
    for (int64_t k = 0; k < n_surf; k++)
        for (int q = 0; q < R.n; q++) {
            const double* xs = surf_x + (size_t)k * P * dim;
            const size_t o = (size_t)k * R.n + q;
            if (x)
                for (int d = 0; d < dim; d++) {
                    double v = 0.;
                    for (int p = 0; p < P; p++) v += xs[p * dim + d] * N[(size_t)q * P + p];
                    x[o * dim + d] = v;
                }
            double nr[3] = {0., 0., 0.};
            surface_normal_raw(dim, P, xs, &dN[(size_t)q * P * ld], nr);
            const double len = (dim == 3) ? std::sqrt(nr[0] * nr[0] + (nr[1] * nr[1] + nr[2] * nr[2])) : std::sqrt(nr[0] * nr[0] + nr[1] * nr[1]);
            if (detg) detg[o] = len;
            if (normal) for (int d = 0; d < dim; d++) normal[o * dim + d] = nr[d] / len;
        }
    return R.n;
}

inline int64_t number_dofs(int64_t n_obj, int dof_size, const uint8_t* status, int64_t init, int64_t* eqn) {
    int64_t c = init;
    for (int64_t k = 0; k < n_obj * dof_size; k++) eqn[k] = (status[k] == ACTIVE) ? c++ : -1;
    return c - init;
}

}  // namespace isl
