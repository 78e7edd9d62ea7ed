// Test infrastructure — NOT product code.
//
// A CPU stand-in for the subset of the C ABI (include/insilico_b200.h) that include/insilico_b200_reference.hpp
// calls, implemented on the CPU oracle (oracle/insilico_oracle.cpp).  It exists so that the binding header can be
// exercised in the `-m "not gpu"` suite: the unmodified reference applications, compiled against the binding and
// linked with THIS file instead of libinsilico_b200.so, must print what they print with base::solver::Eigen3.
// On the GPU the same applications link the real library.  Never shipped, never linked by the product.
#include <cstdint>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <insilico_b200.h>

extern "C" {
void* orc_problem_new();
void orc_problem_free(void*);
void orc_set_mesh(void*, int, int, int, int64_t, const double*, int64_t, const int64_t*);
void orc_set_field(void*, int, int, int, int64_t, const int64_t*, const int64_t*, const uint8_t*, const double*, const double*);
void orc_set_field_constraints(void*, int, int64_t, const int64_t*, const int64_t*, const int64_t*, const double*);
void* orc_system_new(int64_t);
void orc_system_free(void*);
void orc_register_fields(void*, void*, int, int);
int orc_stiffness(void*, void*, int, const double*, int, int, int, int, int);
int orc_residual(void*, void*, int, const double*, int, int, int);
int orc_bodyforce(void*, void*, const double*, int, int);
int orc_bodyforce_sampled(void*, void*, const double*, int, int);
int orc_insert_lhs(void*, const double*, const int64_t*, int, const int64_t*, int);
int orc_stiffness_sampled(void*, void*, int, const double*, int, int, int, int);
int orc_insert_rhs(void*, const double*, const int64_t*, int);
void orc_finish(void*);
int64_t orc_nnz(void*);
void orc_get_csr(void*, int64_t*, int32_t*, double*, double*);
const char* orc_system_error(void*);
double orc_rhs_norm(void*);
int orc_stiffness_aux(void*, void*, int, const double*, int, int, int, int, int);
int orc_residual_aux(void*, void*, int, const double*, int, int, int, int);
int orc_neumann_rows(void*, int, int, int, int, int, int64_t, const double*, const double*, int, const int32_t*, int, const double*);
}

struct isl_engine {
    void* prob = nullptr;
    void* sys = nullptr;
    int64_t n = 0;
    bool finished = false;
    int shape = 0, npe = 0, dim = 0;
    int64_t nElems = 0;
    struct F { int deg = 0, ds = 0; int64_t nObj = 0; std::vector<int64_t> elemDof, eqn; std::vector<uint8_t> status;
               std::vector<double> presc, values;
               std::vector<int64_t> conDof, conPtr, masterEqn; std::vector<double> weight; } f[5];
    std::vector<double> solution;  // rhs after isl_solve_cg
};
static std::string g_err;
// ISL_MOCK_TRACE=1: count the ABI calls and print them at exit (which path did the application take?)
static long g_calls[7] = {0, 0, 0, 0, 0, 0, 0};  // assemble_matrix, assemble_residual, assemble_bodyforce, insert_lhs, insert_rhs, solve_cg, assemble_neumann
static struct TraceAtExit {
    ~TraceAtExit() {
        if (std::getenv("ISL_MOCK_TRACE"))
            std::fprintf(stderr, "[mock abi] assemble_matrix %ld  assemble_residual %ld  assemble_bodyforce %ld  insert_lhs %ld  insert_rhs %ld  solve_cg %ld  assemble_neumann %ld\n",
                         g_calls[0], g_calls[1], g_calls[2], g_calls[3], g_calls[4], g_calls[5], g_calls[6]);
    }
} g_trace;
static int fail(const std::string& m) { g_err = m; return 1; }

extern "C" {
const char* isl_last_error(void) { return g_err.c_str(); }
int isl_engine_create(int, isl_handle* out) { *out = new isl_engine(); (*out)->prob = orc_problem_new(); return 0; }
int isl_engine_destroy(isl_handle h) { delete h; return 0; }
int isl_mesh_set(isl_handle h, int shape, int gdeg, int dim, int64_t nn, const double* x, int64_t ne, const int32_t* conn) {
    // geometry degree 1 only here
    const int npe = (shape == ISL_TRI ? 3 : shape == ISL_QUAD ? 4 : shape == ISL_TET ? 4 : shape == ISL_HEX ? 8 : 2);
    if (gdeg != 1) return fail("mock ABI: geometry degree 1 only");
    std::vector<int64_t> c(conn, conn + ne * npe);
    orc_set_mesh(h->prob, shape, gdeg, dim, nn, x, ne, c.data());
    h->shape = shape; h->npe = npe; h->nElems = ne; h->dim = dim;
    return 0;
}
int isl_mesh_update_coords(isl_handle, const double*) { return fail("mock ABI: isl_mesh_update_coords not provided"); }
static void push_field(isl_handle h, int i) {
    isl_engine::F& f = h->f[i];
    orc_set_field(h->prob, i, f.deg, f.ds, f.nObj, f.elemDof.data(), f.eqn.data(), f.status.data(), f.presc.data(),
                  f.values.data());
    orc_set_field_constraints(h->prob, i, (int64_t)f.conDof.size(), f.conDof.data(), f.conPtr.data(), f.masterEqn.data(),
                              f.weight.data());
}
int isl_field_set(isl_handle h, int i, int deg, int ds, int64_t nObj, const int32_t* ed, const int64_t* eqn,
                  const uint8_t* status, const double* presc, const double* values) {
    isl_engine::F& f = h->f[i];
    f.deg = deg; f.ds = ds; f.nObj = nObj;
    const int64_t ndpe = [&] {  // DoF objects per element of a Lagrange element (base/fe/LagrangeElement.hpp:96-111)
        const int d = deg;
        switch (h->shape) {
            case ISL_TRI: return (int64_t)(d + 1) * (d + 2) / 2;
            case ISL_QUAD: return (int64_t)(d + 1) * (d + 1);
            case ISL_TET: return (int64_t)(d + 1) * (d + 2) * (d + 3) / 6;
            default: return (int64_t)(d + 1) * (d + 1) * (d + 1);
        }
    }();
    f.elemDof.assign(ed, ed + h->nElems * ndpe);
    const size_t n = (size_t)nObj * ds;
    f.eqn.assign(eqn, eqn + n);
    for (size_t k = 0; k < n; k++) if (status[k] != ISL_ACTIVE) f.eqn[k] = -1;
    f.status.assign(status, status + n);
    f.presc.assign(presc, presc + n);
    f.values.assign(values, values + n);
    f.conDof.clear(); f.conPtr.assign(1, 0); f.masterEqn.clear(); f.weight.clear();  // isl_field_set drops constraints
    push_field(h, i);
    return 0;
}
int isl_field_set_constraints(isl_handle h, int i, int64_t nCon, const int64_t* conDof, const int64_t* conPtr,
                              const int64_t* masterEqn, const double* weight) {
    isl_engine::F& f = h->f[i];
    f.conDof.assign(conDof, conDof + nCon);
    f.conPtr.assign(1, 0);
    if (nCon > 0) f.conPtr.assign(conPtr, conPtr + nCon + 1);
    const int64_t nm = nCon > 0 ? conPtr[nCon] : 0;
    f.masterEqn.assign(masterEqn, masterEqn + nm);
    f.weight.assign(weight, weight + nm);
    push_field(h, i);
    return 0;
}
int isl_field_update(isl_handle h, int i, const double* presc, const double* values) {
    isl_engine::F& f = h->f[i];
    const size_t n = (size_t)f.nObj * f.ds;
    if (presc) f.presc.assign(presc, presc + n);
    if (values) f.values.assign(values, values + n);
    push_field(h, i);
    return 0;
}
int isl_system_create(isl_handle h, int64_t n) {
    if (h->sys) orc_system_free(h->sys);
    h->sys = orc_system_new(n);
    h->n = n; h->finished = false; h->solution.clear();
    return 0;
}
int isl_pattern_register(isl_handle h, int t, int c) { orc_register_fields(h->sys, h->prob, t, c); return 0; }
int isl_assemble_matrix(isl_handle h, int kid, const double* p, int q, int t, int c, int incr) {
    g_calls[0]++;
    return orc_stiffness(h->sys, h->prob, kid, p, q, t, c, incr, 1) ? fail(orc_system_error(h->sys)) : 0;
}
int isl_assemble_matrix_aux(isl_handle h, int kid, const double* p, int q, int t, int c, int aux, int incr) {
    if (aux < 0) return isl_assemble_matrix(h, kid, p, q, t, c, incr);
    g_calls[0]++;
    return orc_stiffness_aux(h->sys, h->prob, kid, p, q, t, c, aux, incr) ? fail(orc_system_error(h->sys)) : 0;
}
int isl_assemble_residual_aux(isl_handle h, int kid, const double* p, int q, int t, int c, int aux, double factor) {
    if (aux < 0) return isl_assemble_residual(h, kid, p, q, t, c, factor);
    g_calls[1]++;
    if (factor != -1.0) return fail("mock ABI: residual factor must be -1");
    return orc_residual_aux(h->sys, h->prob, kid, p, q, t, c, aux) ? fail(orc_system_error(h->sys)) : 0;
}
int isl_assemble_matrix_sampled(isl_handle h, int kid, const double* values, int q, int t, int c, int incr) {
    g_calls[0]++;
    return orc_stiffness_sampled(h->sys, h->prob, kid, values, q, t, c, incr) ? fail(orc_system_error(h->sys)) : 0;
}
int isl_assemble_residual(isl_handle h, int kid, const double* p, int q, int t, int c, double factor) {
    g_calls[1]++;
    if (factor != -1.0) return fail("mock ABI: residual factor must be -1");
    return orc_residual(h->sys, h->prob, kid, p, q, t, c) ? fail(orc_system_error(h->sys)) : 0;
}
int isl_assemble_bodyforce(isl_handle h, const double* f, int q, int t) {
    g_calls[2]++;
    return orc_bodyforce(h->sys, h->prob, f, q, t) ? fail(orc_system_error(h->sys)) : 0;
}
int isl_assemble_bodyforce_sampled(isl_handle h, const double* v, int q, int t) {
    g_calls[2]++;
    return orc_bodyforce_sampled(h->sys, h->prob, v, q, t) ? fail(orc_system_error(h->sys)) : 0;
}
int isl_assemble_neumann_rows(isl_handle h, int shape, int gdeg, int64_t nSurf, const double* sx, const double* sp, int q, int feDeg,
                              int ds, const int32_t* rows, int mode, const double* data) {
    g_calls[6]++;
    const int dim = (shape == ISL_TET || shape == ISL_HEX) ? 3 : 2;
    return orc_neumann_rows(h->sys, shape, gdeg, dim, feDeg, ds, nSurf, sx, sp, q, rows, mode, data);
}
int isl_insert_lhs(isl_handle h, const double* m, const int64_t* r, int nr, const int64_t* c, int nc) {
    g_calls[3]++;
    return orc_insert_lhs(h->sys, m, r, nr, c, nc) ? fail(orc_system_error(h->sys)) : 0;
}
int isl_insert_rhs(isl_handle h, const double* v, const int64_t* r, int nr) { g_calls[4]++; return orc_insert_rhs(h->sys, v, r, nr); }
int isl_finish(isl_handle h, int64_t* n, int64_t* nnz) {
    if (!h->finished) { orc_finish(h->sys); h->finished = true; }
    if (n) *n = h->n;
    if (nnz) *nnz = orc_nnz(h->sys);
    return 0;
}
int isl_get_csr(isl_handle h, int64_t* rowptr, int32_t* col, double* val, double* rhs) {
    if (!h->finished) { orc_finish(h->sys); h->finished = true; }
    orc_get_csr(h->sys, rowptr, col, val, rhs);
    if (rhs && !h->solution.empty()) std::memcpy(rhs, h->solution.data(), h->solution.size() * sizeof(double));
    return 0;
}
int isl_get_device_csr(isl_handle, int64_t**, int32_t**, double**, double**) { return fail("mock ABI: no device"); }
int isl_rhs_value(isl_handle h, int64_t i, double* v) {
    std::vector<double> b((size_t)h->n);
    orc_get_csr(h->sys, nullptr, nullptr, nullptr, b.data());
    *v = b[(size_t)i];
    return 0;
}
int isl_rhs_norm(isl_handle h, double* v) {
    if (!h->solution.empty()) { double a = 0.; for (double x : h->solution) a += x * x; *v = std::sqrt(a) / (double)h->n; return 0; }
    *v = orc_rhs_norm(h->sys);
    return 0;
}
// Jacobi-preconditioned CG as in Eigen 3.2's ConjugateGradient.h (what isl_solve_cg does on the device)
int isl_solve_cg(isl_handle h, double tol, int64_t maxIter, int64_t* iterations, double* error) {
    g_calls[5]++;
    int64_t n = 0, nnz = 0;
    isl_finish(h, &n, &nnz);
    std::vector<int64_t> rp(n + 1); std::vector<int32_t> col(nnz); std::vector<double> val(nnz), b(n);
    orc_get_csr(h->sys, rp.data(), col.data(), val.data(), b.data());
    if (tol <= 0.) tol = 2.220446049250313e-16;
    if (maxIter <= 0) maxIter = 2 * n;
    std::vector<double> x(n, 0.), r(b), p(n), Ap(n), dinv(n, 1.);
    for (int64_t i = 0; i < n; i++)
        for (int64_t k = rp[i]; k < rp[i + 1]; k++) if (col[k] == i && val[k] != 0.) dinv[i] = 1. / val[k];
    double bb = 0., rr = 0., rz = 0.;
    for (int64_t i = 0; i < n; i++) { bb += b[i] * b[i]; p[i] = dinv[i] * r[i]; rz += r[i] * p[i]; }
    rr = bb;
    int64_t it = 0;
    if (bb > 0.)
        while (rr >= tol * tol * bb && it < maxIter) {
            double pAp = 0.;
            for (int64_t i = 0; i < n; i++) { double s = 0.; for (int64_t k = rp[i]; k < rp[i + 1]; k++) s += val[k] * p[col[k]]; Ap[i] = s; pAp += p[i] * s; }
            const double alpha = rz / pAp;
            rr = 0.; double rzn = 0.;
            for (int64_t i = 0; i < n; i++) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; rr += r[i] * r[i]; rzn += r[i] * dinv[i] * r[i]; }
            if (rr < tol * tol * bb) break;
            const double beta = rzn / rz; rz = rzn;
            for (int64_t i = 0; i < n; i++) p[i] = dinv[i] * r[i] + beta * p[i];
            it++;
        }
    h->solution = x;
    if (iterations) *iterations = it;
    if (error) *error = bb > 0. ? std::sqrt(rr / bb) : 0.;
    return 0;
}
}
