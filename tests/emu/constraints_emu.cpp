// Test infrastructure: host replay of the engine's treatment of slaves of master DoFs (general linear constraints).
// Uses the SAME index routines as the CUDA kernels (insilico_b200/csrc/isl_constraints.hpp: isl_scatter_constrained,
// isl_scatter_force_to_masters, isl_effective_ids, isl_find_in_row) with plain additions instead of atomics, and
// replays around them what isl_engine.cu does: pattern keys (ACTIVE x ACTIVE per element + effective rows x columns of
// elements holding a slave), sort/unique into CSR, slot map, scatter_entry, the force scatter of k_force.
// Local matrices / force vectors are handed in (random numbers in the test), so only the scatter logic is under test.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../insilico_b200/csrc/isl_constraints.hpp"

namespace {
struct HostAdd {
    static void add(double* target, double value) { *target += value; }
};
struct FieldView {
    const int32_t* elem_dof; int ndpe, ds;
    const int32_t* eqn; const uint8_t* status; const double* presc; const double* values;
    const int32_t* cptr; const int32_t* cm; const double* cw;  // cptr may be null
};
}  // namespace

extern "C" {

// returns nnz; out arrays sized by the caller (rowptr[n_eqn+1], col/val[cap], rhs[n_eqn]); -1 if cap is too small
int64_t emu_constraints(int64_t n_elems, int64_t n_eqn, const int32_t* ed_t, int ndpe_t, int ds_t, const int32_t* eqn_t,
                        const uint8_t* st_t, const int32_t* cptr_t, const int32_t* cm_t, const double* cw_t,
                        const int32_t* ed_c, int ndpe_c, int ds_c, const int32_t* eqn_c, const uint8_t* st_c,
                        const double* presc_c, const double* val_c, const int32_t* cptr_c, const int32_t* cm_c,
                        const double* cw_c, int incremental, const double* Kloc, const double* floc, int64_t cap,
                        int64_t* rowptr, int32_t* col, double* val, double* rhs) {
    const FieldView T{ed_t, ndpe_t, ds_t, eqn_t, st_t, nullptr, nullptr, cptr_t, cm_t, cw_t};
    const FieldView C{ed_c, ndpe_c, ds_c, eqn_c, st_c, presc_c, val_c, cptr_c, cm_c, cw_c};
    const int nr = ndpe_t * ds_t, ncl = ndpe_c * ds_c;
    // ---- pattern (build_pattern): k_make_keys + host extras
    std::vector<uint64_t> keys;
    for (int64_t e = 0; e < n_elems; e++)
        for (int i = 0; i < nr; i++)
            for (int j = 0; j < ncl; j++) {
                const int32_t r = T.eqn[(size_t)T.elem_dof[e * T.ndpe + i / T.ds] * T.ds + i % T.ds];
                const int32_t c = C.eqn[(size_t)C.elem_dof[e * C.ndpe + j / C.ds] * C.ds + j % C.ds];
                if (r >= 0 && c >= 0) keys.push_back(((uint64_t)(uint32_t)r << 32) | (uint32_t)c);
            }
    if (T.cptr || C.cptr) {
        std::vector<int32_t> er, ec;
        for (int64_t e = 0; e < n_elems; e++) {
            const bool sr = isl_effective_ids(T.elem_dof, T.ndpe, T.ds, T.eqn, T.cptr, T.cm, e, er);
            const bool sc = isl_effective_ids(C.elem_dof, C.ndpe, C.ds, C.eqn, C.cptr, C.cm, e, ec);
            if (!sr && !sc) continue;
            for (int32_t r : er) for (int32_t c : ec) keys.push_back(((uint64_t)(uint32_t)r << 32) | (uint32_t)c);
        }
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    const int64_t nnz = (int64_t)keys.size();
    if (nnz > cap) return -1;
    std::fill(rowptr, rowptr + n_eqn + 1, 0);
    for (int64_t k = 0; k < nnz; k++) { rowptr[(keys[k] >> 32) + 1]++; col[k] = (int32_t)(keys[k] & 0xffffffffu); val[k] = 0.; }
    for (int64_t r = 0; r < n_eqn; r++) rowptr[r + 1] += rowptr[r];
    std::fill(rhs, rhs + n_eqn, 0.);
    const IslMasters mt{T.cptr, T.cm, T.cw}, mc{C.cptr, C.cm, C.cw};
    // ---- matrix entries (scatter_entry)
    for (int64_t e = 0; e < n_elems; e++)
        for (int i = 0; i < nr; i++)
            for (int j = 0; j < ncl; j++) {
                const double v = Kloc[((size_t)e * nr + i) * ncl + j];
                const size_t kr = (size_t)T.elem_dof[e * T.ndpe + i / T.ds] * T.ds + i % T.ds;
                const size_t k = (size_t)C.elem_dof[e * C.ndpe + j / C.ds] * C.ds + j % C.ds;
                const int32_t r = T.eqn[kr], c = C.eqn[k];
                const int64_t slot = (r < 0 || c < 0) ? -1 : isl_find_in_row(rowptr, col, r, c);  // k_slotmap
                if (slot >= 0) { val[slot] += v; continue; }
                const bool c_con = (C.status[k] == 1);
                const double g = c_con ? (incremental ? C.presc[k] - C.values[k] : C.presc[k]) : 0.;
                if (T.cptr != nullptr || C.cptr != nullptr) {
                    isl_scatter_constrained<HostAdd>(mt, mc, rowptr, col, val, rhs, kr, r, k, c, c_con, g, v);
                    continue;
                }
                if (r < 0) continue;
                if (c_con) rhs[r] += -(g * v);
            }
    // ---- forces (k_force)
    if (floc)
        for (int64_t e = 0; e < n_elems; e++)
            for (int i = 0; i < nr; i++) {
                const size_t kr = (size_t)T.elem_dof[e * T.ndpe + i / T.ds] * T.ds + i % T.ds;
                const int32_t r = T.eqn[kr];
                const double f = floc[(size_t)e * nr + i];
                if (r >= 0) rhs[r] += f;
                else if (T.cptr != nullptr) isl_scatter_force_to_masters<HostAdd>(mt, rhs, kr, f);
            }
    return nnz;
}
}
