// Host-side reference-element tables of the B200 assembly engine: quadrature rules, Lagrange bases in
// the reference's hierarchic local ordering, and n-face topology.  These are evaluated once per
// (rule, basis) pair and shipped to the device as flat constant tables.
//
// Reference behaviour followed (paths relative to the reference root):
//   base/Quadrature.hpp:28-79, base/quad/GaussLegendre.hpp:36,116-218, base/quad/TensorProduct.hpp:122-177,
//   base/quad/GaussTetrahedron.hpp:93-176, base/quad/GaussTriangle.hpp:97-180,
//   base/sfun/Lagrange1D.ipp, base/sfun/TensorProduct.hpp, base/sfun/LagrangeTetrahedron.ipp,
//   base/sfun/LagrangeTriangle.ipp, base/mesh/HierarchicOrder.hpp, base/mesh/ElementFaces.hpp.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace isl {

enum { POINT = 0, LINE = 1, TRI = 2, QUAD = 3, TET = 4, HEX = 5 };
enum { VERTEX = 0, EDGE = 1, FACE = 2, CELL = 3 };
enum { ACTIVE = 0, CONSTRAINED = 1, INACTIVE = 2 };

inline int shape_dim(int s) {
    if (s == LINE) return 1;
    if (s == TRI || s == QUAD) return 2;
    if (s == TET || s == HEX) return 3;
    throw std::runtime_error("unsupported shape " + std::to_string(s));
}
inline bool is_cube(int s) { return s == LINE || s == QUAD || s == HEX; }

// number of vertices / edges / faces of a shape
inline int n_subfaces(int s, int nf) {
    static const int t[6][4] = {{1, 0, 0, 0}, {2, 1, 0, 0}, {3, 3, 1, 0}, {4, 4, 1, 0}, {4, 6, 4, 1}, {8, 12, 6, 1}};
    return t[s][nf];
}

// ---------------------------------------------------------------------------------------------------
// topology tables (vertex numbers in hierarchic order), ElementFaces.hpp:224-434
struct Topology {
    int n_vert, n_edge, n_face, face_nv;
    int edge[12][2];
    int face[6][4];
    int face_edge[6][4];  // FaceEdges index, ElementFaces.hpp:529-600
    int face_edge_sign[6][4];
};

inline const Topology& topology(int s) {
    static const Topology tri = {3, 3, 1, 3, {{0, 1}, {1, 2}, {2, 0}}, {{0, 1, 2, -1}}, {{0, 1, 2, -1}}, {{1, 1, 1, 0}}};
    static const Topology quad = {4, 4, 1, 4, {{0, 1}, {1, 2}, {2, 3}, {3, 0}}, {{0, 1, 2, 3}}, {{0, 1, 2, 3}},
                                  {{1, 1, 1, 1}}};
    static const Topology tet = {4, 6, 4, 3,
                                 {{0, 1}, {1, 2}, {2, 0}, {3, 0}, {3, 1}, {3, 2}},
                                 {{0, 2, 1, -1}, {0, 1, 3, -1}, {1, 2, 3, -1}, {2, 0, 3, -1}},
                                 {{2, 1, 0, -1}, {0, 4, 3, -1}, {1, 5, 4, -1}, {2, 3, 5, -1}},
                                 {{-1, -1, -1, 0}, {1, 1, -1, 0}, {1, 1, -1, 0}, {1, 1, -1, 0}}};
    static const Topology hex = {
        8, 12, 6, 4,
        {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}},
        {{0, 3, 2, 1}, {4, 5, 6, 7}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}},
        {{3, 2, 1, 0}, {4, 5, 6, 7}, {0, 9, 4, 8}, {1, 10, 5, 9}, {2, 11, 6, 10}, {3, 8, 7, 11}},
        {{-1, -1, -1, -1}, {1, 1, 1, 1}, {1, 1, -1, -1}, {1, 1, -1, -1}, {1, 1, -1, -1}, {1, 1, -1, -1}}};
    static const Topology line = {2, 1, 0, 0, {{0, 1}}, {}, {}, {}};
    switch (s) {
        case LINE: return line;
        case TRI: return tri;
        case QUAD: return quad;
        case TET: return tet;
        case HEX: return hex;
    }
    throw std::runtime_error("unsupported shape");
}

// DoFs per n-face of a Lagrange element and their position in the element's DoF array
// (fe/LagrangeElement.hpp:25-115, fe/Policies.hpp:38-77)
struct FELayout {
    int per[4], count[4], begin[4], total;
};
inline FELayout fe_layout(int shape, int deg) {
    if (deg < 1) throw std::runtime_error("Lagrange degree >= 1 required");
    const int dim = shape_dim(shape);
    auto pw = [](int m, int n) { int r = 1; while (n-- > 0) r *= m; return r; };
    auto binom = [](int n, int k) { if (k > n || k < 0) return 0; long r = 1; for (int i = 1; i <= k; i++) r = r * (n - k + i) / i; return (int)r; };
    FELayout L;
    L.per[VERTEX] = 1;
    for (int nf = 1; nf <= 3; nf++)
        L.per[nf] = (dim >= nf) ? (is_cube(shape) ? pw(deg - 1, nf) : binom(deg - 1, nf)) : 0;
    L.count[VERTEX] = n_subfaces(shape, VERTEX);
    L.count[EDGE] = n_subfaces(shape, EDGE);
    L.count[FACE] = n_subfaces(shape, FACE);
    L.count[CELL] = dim == 3 ? 1 : 0;
    int pos = 0;
    for (int nf = 0; nf < 4; nf++) { L.begin[nf] = pos; pos += L.per[nf] * L.count[nf]; }
    L.total = pos;
    return L;
}

// ---------------------------------------------------------------------------------------------------
// hierarchic position of every lexicographic tensor node (i,j,k) in {0..K}^dim: vertices first, then edge
// interiors walked from the edge's first to its second vertex, then face interiors, then the cell interior.
inline std::vector<int> lexi_to_hier(int shape, int K) {
    const int dim = shape_dim(shape);
    if (!is_cube(shape)) throw std::runtime_error("lexi_to_hier: hypercubes only");
    const int n1 = K + 1;
    int total = 1; for (int d = 0; d < dim; d++) total *= n1;
    std::vector<int> H(total, -1);
    auto lin = [&](const int* c) { int r = 0; for (int d = dim - 1; d >= 0; d--) r = r * n1 + c[d]; return r; };
    // corner coordinates of the hierarchic vertices (counter-clockwise bottom, then top)
    static const int corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    const int nv = 1 << dim;
    const int vmap1[2] = {0, 1};
    int next = 0;
    auto vcoord = [&](int v, int* c) {
        if (dim == 1) { c[0] = vmap1[v] * K; return; }
        for (int d = 0; d < dim; d++) c[d] = corner[v][d] * K;
    };
    for (int v = 0; v < nv; v++) { int c[3]; vcoord(v, c); H[lin(c)] = next++; }
    if (dim == 1) { for (int i = 1; i < K; i++) { int c[3] = {i, 0, 0}; H[lin(c)] = next++; } return H; }
    const Topology& T = topology(shape);
    for (int e = 0; e < T.n_edge; e++) {
        int a[3], b[3]; vcoord(T.edge[e][0], a); vcoord(T.edge[e][1], b);
        for (int n = 1; n < K; n++) {
            int c[3] = {0, 0, 0};
            for (int d = 0; d < dim; d++) c[d] = a[d] + (b[d] - a[d]) / K * n;
            H[lin(c)] = next++;
        }
    }
    // faces are spanned from their first vertex towards two neighbours (HierarchicOrder.hpp:383-386)
    static const int span3[6][3] = {{0, 1, 3}, {4, 5, 7}, {0, 1, 4}, {1, 2, 5}, {2, 3, 6}, {3, 0, 7}};
    static const int span2[1][3] = {{0, 1, 3}};
    const int nfaces = dim == 2 ? 1 : 6;
    for (int f = 0; f < nfaces; f++) {
        const int* sp = dim == 2 ? span2[f] : span3[f];
        int o[3], u[3], w[3]; vcoord(sp[0], o); vcoord(sp[1], u); vcoord(sp[2], w);
        for (int n2 = 1; n2 < K; n2++)
            for (int n1i = 1; n1i < K; n1i++) {
                int c[3] = {0, 0, 0};
                for (int d = 0; d < dim; d++) c[d] = o[d] + (u[d] - o[d]) / K * n1i + (w[d] - o[d]) / K * n2;
                H[lin(c)] = next++;
            }
    }
    if (dim == 3)
        for (int k = 1; k < K; k++) for (int j = 1; j < K; j++) for (int i = 1; i < K; i++) { int c[3] = {i, j, k}; H[lin(c)] = next++; }
    return H;
}

// ---------------------------------------------------------------------------------------------------
// Lagrange bases
struct Basis {
    int shape, deg, dim, nfun;
    std::vector<int> H;  // lexicographic -> hierarchic (cubes)

    Basis(int s, int d) : shape(s), deg(d), dim(shape_dim(s)) {
        if (is_cube(s)) { H = lexi_to_hier(s, d); nfun = (int)H.size(); }
        else {
            if (d < 1 || d > 2) throw std::runtime_error("simplex Lagrange degree 1 or 2 supported");
            nfun = (s == TRI) ? (d == 1 ? 3 : 6) : (d == 1 ? 4 : 10);
        }
        if (is_cube(s) && (d < 1 || d > 3)) throw std::runtime_error("cube Lagrange degree 1..3 supported");
    }

    static void poly1d(int deg, double x, double* v, double* g) {
        switch (deg) {
            case 1: v[0] = 1. - x; v[1] = x; g[0] = -1.; g[1] = 1.; break;
            case 2:
                v[0] = (1. - x) * (1. - 2. * x); v[1] = 4. * x * (1. - x); v[2] = x * (2. * x - 1.);
                g[0] = 4. * x - 3.; g[1] = 4. - 8. * x; g[2] = 4. * x - 1.;
                break;
            case 3: {
                const double a = 1. - x, b = x;
                v[0] = 0.5 * a * (3. * a - 1.) * (3. * a - 2.); v[1] = 4.5 * a * (3. * a - 1.) * b;
                v[2] = 4.5 * b * (3. * b - 1.) * a; v[3] = 0.5 * b * (3. * b - 1.) * (3. * b - 2.);
                g[0] = -0.5 * ((3. * a - 1.) * (3. * a - 2.) + 3. * a * (6. * a - 3.));
                g[1] = -4.5 * (b * (6. * a - 1.) - a * (3. * a - 1.));
                g[2] = 4.5 * (a * (6. * b - 1.) - b * (3. * b - 1.));
                g[3] = 0.5 * ((3. * b - 1.) * (3. * b - 2.) + 3. * b * (6. * b - 3.));
            } break;
        }
    }

    // fun[nfun], grad[nfun*dim] (either may be null)
    void eval(const double* xi, double* fun, double* grad) const {
        if (is_cube(shape)) {
            double v[3][4], g[3][4];
            for (int d = 0; d < dim; d++) poly1d(deg, xi[d], v[d], g[d]);
            const int n1 = deg + 1;
            for (int n = 0; n < nfun; n++) {
                int idx[3] = {0, 0, 0}, r = n;
                for (int d = 0; d < dim; d++) { idx[d] = r % n1; r /= n1; }
                const int h = H[n];
                if (dim == 1) {
                    if (fun) fun[h] = v[0][idx[0]];
                    if (grad) grad[h] = g[0][idx[0]];
                } else if (dim == 2) {
                    if (fun) fun[h] = v[0][idx[0]] * v[1][idx[1]];
                    if (grad) { grad[h * 2] = v[1][idx[1]] * g[0][idx[0]]; grad[h * 2 + 1] = g[1][idx[1]] * v[0][idx[0]]; }
                } else {
                    // association as in sfun/TensorProduct.hpp: N_z*(N_y*dN_x), N_z*(dN_y*N_x), dN_z*(N_x*N_y)
                    if (fun) fun[h] = (v[0][idx[0]] * v[1][idx[1]]) * v[2][idx[2]];
                    if (grad) {
                        grad[h * 3 + 0] = v[2][idx[2]] * (v[1][idx[1]] * g[0][idx[0]]);
                        grad[h * 3 + 1] = v[2][idx[2]] * (g[1][idx[1]] * v[0][idx[0]]);
                        grad[h * 3 + 2] = g[2][idx[2]] * (v[0][idx[0]] * v[1][idx[1]]);
                    }
                }
            }
            return;
        }
        // simplices: barycentric formulation, vertices then edge midpoints in ElementFaces edge order
        const int nv = dim + 1;
        double z[4], dz[4][3];
        z[0] = 1.; for (int d = 0; d < dim; d++) z[0] -= xi[d];
        for (int d = 0; d < dim; d++) { z[d + 1] = xi[d]; dz[0][d] = -1.; for (int a = 1; a < nv; a++) dz[a][d] = (a == d + 1) ? 1. : 0.; }
        if (deg == 1) {
            for (int a = 0; a < nv; a++) { if (fun) fun[a] = z[a]; if (grad) for (int d = 0; d < dim; d++) grad[a * dim + d] = dz[a][d]; }
            return;
        }
        for (int a = 0; a < nv; a++) {
            if (fun) fun[a] = z[a] * (2. * z[a] - 1.);
            if (grad) for (int d = 0; d < dim; d++) grad[a * dim + d] = (4. * z[a] - 1.) * dz[a][d];
        }
        // edge function pairs (a,b): TRI (1,0)->listed as 4 x0 z0 ... ; use the reference's pairs
        static const int tri_pairs[3][2] = {{1, 0}, {1, 2}, {2, 0}};
        static const int tet_pairs[6][2] = {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}};
        const int ne = dim == 2 ? 3 : 6;
        for (int e = 0; e < ne; e++) {
            const int a = dim == 2 ? tri_pairs[e][0] : tet_pairs[e][0], b = dim == 2 ? tri_pairs[e][1] : tet_pairs[e][1];
            if (fun) fun[nv + e] = 4. * z[a] * z[b];
            if (grad) for (int d = 0; d < dim; d++) grad[(nv + e) * dim + d] = 4. * (z[a] * dz[b][d] + dz[a][d] * z[b]);
        }
    }

    // support points pts[nfun*dim]
    void support(double* pts) const {
        if (is_cube(shape)) {
            const int n1 = deg + 1;
            for (int n = 0; n < nfun; n++) {
                int r = n;
                for (int d = 0; d < dim; d++) { pts[H[n] * dim + d] = double(r % n1) / double(deg); r /= n1; }
            }
            return;
        }
        const int nv = dim + 1;
        for (int a = 0; a < nv; a++) for (int d = 0; d < dim; d++) pts[a * dim + d] = (a == d + 1) ? 1. : 0.;
        if (deg == 2) {
            const Topology& T = topology(shape);
            for (int e = 0; e < T.n_edge; e++)
                for (int d = 0; d < dim; d++)
                    pts[(nv + e) * dim + d] = 0.5 * (pts[T.edge[e][0] * dim + d] + pts[T.edge[e][1] * dim + d]);
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// quadrature rules: w[n], p[n*dim]
struct Rule {
    int n = 0, dim = 0;
    std::vector<double> w, p;
};

inline void gauss_legendre_01(int npts, const double*& w, const double*& x) {
    // (weight, point) literals of the reference, base/quad/GaussLegendre.hpp:116-218
    static const double W1[] = {1};
    static const double X1[] = {0.5};
    static const double W2[] = {0.5, 0.5};
    static const double X2[] = {0.788675134594813, 0.211324865405187};
    static const double W3[] = {0.277777777777777, 0.444444444444444, 0.277777777777777};
    static const double X3[] = {0.887298334620741, 0.5, 0.112701665379259};
    static const double W4[] = {0.173927422568727, 0.326072577431273, 0.326072577431273, 0.173927422568727};
    static const double X4[] = {0.930568155797026, 0.669990521792428, 0.330009478207572, 0.069431844202974};
    static const double W5[] = {0.118463442528095, 0.239314335249683, 0.284444444444444, 0.239314335249683,
                                0.118463442528095};
    static const double X5[] = {0.953089922969332, 0.769234655052841, 0.5, 0.230765344947159, 0.046910077030668};
    static const double W6[] = {0.085662246189585, 0.180380786524069, 0.233956967286345,
                                0.233956967286345, 0.180380786524069, 0.085662246189585};
    static const double X6[] = {0.966234757101576, 0.830604693233132, 0.619309593041598,
                                0.380690406958402, 0.169395306766868, 0.033765242898424};
    switch (npts) {
        case 1: w = W1; x = X1; return;
        case 2: w = W2; x = X2; return;
        case 3: w = W3; x = X3; return;
        case 4: w = W4; x = X4; return;
        case 5: w = W5; x = X5; return;
        case 6: w = W6; x = X6; return;
    }
    throw std::runtime_error("Gauss-Legendre rule with " + std::to_string(npts) + " points not tabulated");
}

inline Rule make_rule(int shape, int degree) {
    Rule R;
    R.dim = shape_dim(shape);
    if (degree < 1) throw std::runtime_error("quadrature degree must be positive");
    if (is_cube(shape)) {
        const int n1 = (degree + 2) / 2;  // GaussLegendre.hpp:36
        const double *w1, *x1;
        gauss_legendre_01(n1, w1, x1);
        int n = 1; for (int d = 0; d < R.dim; d++) n *= n1;
        R.n = n; R.w.resize(n); R.p.resize((size_t)n * R.dim);
        for (int q = 0; q < n; q++) {  // x fastest; weight (w_x*w_y)*w_z
            int r = q; double w = 1.;
            for (int d = 0; d < R.dim; d++) { const int i = r % n1; r /= n1; w = (d == 0) ? w1[i] : w * w1[i]; R.p[(size_t)q * R.dim + d] = x1[i]; }
            R.w[q] = w;
        }
        return R;
    }
    auto add = [&](double w, double a, double b, double c = 0.) {
        R.w.push_back(w); R.p.push_back(a); R.p.push_back(b); if (R.dim == 3) R.p.push_back(c); R.n++;
    };
    // symmetric orbits written out in the reference's point order
    auto orbit4 = [&](double w, double x, double y) { add(w, x, x, x); add(w, y, x, x); add(w, x, y, x); add(w, x, x, y); };
    auto orbit6 = [&](double w, double x, double y) { add(w, y, y, x); add(w, y, x, x); add(w, x, y, x); add(w, x, x, y); add(w, y, x, y); add(w, x, y, y); };
    if (shape == TET) {  // GaussTetrahedron.hpp:93-176
        switch (degree) {
            case 1: add(0.166666666666666, 0.25, 0.25, 0.25); break;
            case 2: orbit4(0.04166666666666666667, 0.13819660112501051518, 0.58541019662496845446); break;
            case 3: add(-0.13333333333333333, 0.25, 0.25, 0.25); orbit4(0.075, 0.1666666666666667, 0.5); break;
            case 4:
                add(-0.01315555555555555556, 0.25, 0.25, 0.25);
                orbit4(0.0076222222222222, 0.071428571428571, 0.785714285714286);
                orbit6(0.024888888888889, 0.100596423833201, 0.399403576166799);
                break;
            case 5:
                add(0.030283678097089, 0.25, 0.25, 0.25);
                orbit4(0.006026785714286, 0.333333333333333, 0.0);
                orbit4(0.011645249086029, 0.090909090909091, 0.727272727272727);
                orbit6(0.010949141561386, 0.066550153573664, 0.433449846426336);
                break;
            default: throw std::runtime_error("tetrahedron rule degree 1..5 supported");
        }
        return R;
    }
    if (shape == TRI) {  // GaussTriangle.hpp:97-180
        switch (degree) {
            case 1: add(0.5, 0.333333333333333, 0.333333333333333); break;
            case 2:
                add(0.166666666666666, 0.666666666666667, 0.166666666666667);
                add(0.166666666666666, 0.166666666666667, 0.666666666666667);
                add(0.166666666666666, 0.166666666666667, 0.166666666666667);
                break;
            case 3:
                add(-0.28125, 0.333333333333333, 0.333333333333333);
                add(0.260416666666667, 0.6, 0.2); add(0.260416666666667, 0.2, 0.6); add(0.260416666666667, 0.2, 0.2);
                break;
            case 4:
                add(0.111690794839005, 0.10810301816807, 0.445948490915965);
                add(0.054975871827661, 0.816847572980459, 0.091576213509771);
                add(0.111690794839005, 0.445948490915965, 0.10810301816807);
                add(0.111690794839005, 0.445948490915965, 0.445948490915965);
                add(0.054975871827661, 0.091576213509771, 0.816847572980459);
                add(0.054975871827661, 0.091576213509771, 0.091576213509771);
                break;
            case 5:
                add(0.1125, 0.333333333333333, 0.333333333333333);
                add(0.066197076394253, 0.05971587178977, 0.470142064105115);
                add(0.0629695902724135, 0.797426985353087, 0.101286507323456);
                add(0.066197076394253, 0.470142064105115, 0.05971587178977);
                add(0.066197076394253, 0.470142064105115, 0.470142064105115);
                add(0.0629695902724135, 0.101286507323456, 0.797426985353087);
                add(0.0629695902724135, 0.101286507323456, 0.101286507323456);
                break;
            default: throw std::runtime_error("triangle rule degree 1..5 supported");
        }
        return R;
    }
    throw std::runtime_error("quadrature: unsupported shape");
}

}  // namespace isl
