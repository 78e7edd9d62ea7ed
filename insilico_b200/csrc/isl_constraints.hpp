// =============================================================================
// isl_constraints.hpp -- slaves of master DoFs (general linear constraints, base/dof/Constraint.hpp:57-140):
// the index logic shared by the CUDA kernels (isl_engine.cu) and the host replay under tests/emu/.
//
// Reference semantics (base/asmb/assembleMatrix.hpp:56-130,212-338, assembleForces.hpp:58-139): with the row targets of
// a local row r = {(eqn_r, 1)} if ACTIVE, {(master, weight)...} if CONSTRAINED, and the same for a local column c,
//     A[rt, ct]  += w_r * w_c * K(r,c)      for every row target and every column target
//     rhs[rt]    -= g_c * w_r * K(r,c)      for a CONSTRAINED column with prescribed part g_c (also when it has masters)
//     rhs[rt]    += w_r * f(r)              for forces
// =============================================================================
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#if defined(__CUDACC__)
#define ISL_HD __host__ __device__
#else
#define ISL_HD
#endif

// dense constraint table of one field: masters of DoF component k are [cptr[k], cptr[k+1]) (equation numbers cm,
// weights cw); cptr == nullptr when the field has no slaves
struct IslMasters {
    const int32_t* cptr;
    const int32_t* cm;
    const double* cw;
};

ISL_HD inline int64_t isl_find_in_row(const int64_t* rowptr, const int32_t* col, int32_t r, int32_t c) {
    int64_t lo = rowptr[r], hi = rowptr[r + 1];
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (col[mid] < c) lo = mid + 1; else hi = mid; }
    return (lo < rowptr[r + 1] && col[lo] == c) ? lo : -1;
}

// One local entry v = K(r,c) whose row and/or column is not ACTIVE.  kr / k: DoF-component indices of row / column,
// r / c: their equation numbers (< 0 when not ACTIVE), c_con: column CONSTRAINED, g: its prescribed part.
// ADD::add(double* target, double value) accumulates (atomicAdd on the device).
template <class ADD>
ISL_HD inline void isl_scatter_constrained(const IslMasters& mt, const IslMasters& mc, const int64_t* rowptr, const int32_t* col,
                                           double* val, double* rhs, size_t kr, int32_t r, size_t k, int32_t c, bool c_con,
                                           double g, double v) {
    int rb = 0, re = 0;
    if (r < 0) {
        if (mt.cptr == nullptr) return;
        rb = mt.cptr[kr]; re = mt.cptr[kr + 1];
        if (rb == re) return;  // CONSTRAINED without masters or INACTIVE: the row contributes nothing
    }
    if (c < 0 && !c_con) return;  // INACTIVE column
    int cb = 0, ce = 0;
    if (c < 0 && mc.cptr != nullptr) { cb = mc.cptr[k]; ce = mc.cptr[k + 1]; }
    const int nrt = (r >= 0) ? 1 : (re - rb);
    for (int a = 0; a < nrt; a++) {
        const int32_t rt = (r >= 0) ? r : mt.cm[rb + a];
        const double wr = (r >= 0) ? 1.0 : mt.cw[rb + a];
        if (c >= 0) {
            const int64_t pos = isl_find_in_row(rowptr, col, rt, c);
            if (pos >= 0) ADD::add(val + pos, wr * v);
        } else {
            ADD::add(rhs + rt, -(g * wr * v));
            for (int b = cb; b < ce; b++) {
                const int64_t pos = isl_find_in_row(rowptr, col, rt, mc.cm[b]);
                if (pos >= 0) ADD::add(val + pos, wr * mc.cw[b] * v);
            }
        }
    }
}

// force entry f of a local row that is a slave: weight * f to every master (asmb/assembleForces.hpp:118-131)
template <class ADD>
ISL_HD inline void isl_scatter_force_to_masters(const IslMasters& mt, double* rhs, size_t kr, double f) {
    for (int b = mt.cptr[kr]; b < mt.cptr[kr + 1]; b++) ADD::add(rhs + mt.cm[b], mt.cw[b] * f);
}

// effective equation numbers of element e for the sparsity pattern: ACTIVE ids and the masters of its slaves
// (solver/TripletContainer.hpp:229-262); returns whether the element holds a slave.  Host only.
inline bool isl_effective_ids(const int32_t* elem_dof, int ndpe, int ds, const int32_t* eqn, const int32_t* cptr,
                              const int32_t* cm, int64_t e, std::vector<int32_t>& eff) {
    eff.clear();
    bool slave = false;
    for (int d = 0; d < ndpe; d++)
        for (int s = 0; s < ds; s++) {
            const size_t k = (size_t)elem_dof[(size_t)e * ndpe + d] * ds + s;
            if (eqn[k] >= 0) eff.push_back(eqn[k]);
            else if (cptr != nullptr)
                for (int32_t b = cptr[k]; b < cptr[k + 1]; b++) { eff.push_back(cm[b]); slave = true; }
        }
    return slave;
}
