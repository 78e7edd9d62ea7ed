// Host emulation of the row-gather kernel (insilico_b200/csrc/isl_rowgather.cuh) -- TEST INFRASTRUCTURE ONLY.
// Compiles the kernel's per-thread routines (rg_instance, rg_row_affine, rg_add_slot_general, rg_row_tables, the segment arithmetic of the write-out) and the host preprocessing
// (isl_patch_host.hpp) with g++ and replays k_row_meta / k_q1hex_rows_affine patch by patch, thread by thread, including
// the per-warp staging (segments of consecutive rows, 16-byte phase of the bulk copies), so that the algorithm, its tables and its
// index arithmetic can be checked against the oracle on a machine without a GPU (tests/test_rowgather_emu.py).
// Nothing in the product links or loads this file.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/insilico_b200.h"
#include "../../insilico_b200/csrc/isl_tables.hpp"

namespace {
constexpr int sym_idx(int a, int b) {
    return (a < b) ? (a * 8 - (a * (a - 1)) / 2 + (b - a)) : (b * 8 - (b * (b - 1)) / 2 + (a - b));
}
#include "../../insilico_b200/csrc/isl_patch_host.hpp"
#include "../../insilico_b200/csrc/isl_rowgather.cuh"

double g_w0 = 0.;
double g_dN[8 * 8 * 3], g_Nq[64], g_w[8];
// local matrix and body-force integrals of a general trilinear element in plain quadrature order (stands in for the
// device's q1_K_fast, which the gpu tests cover): K_ab = sum_q factor w_q det J_q grad N_a . grad N_b
void host_K(const double (&X)[8][3], double factor, double (&K)[36], double (&bf)[8]) {
    for (int k = 0; k < 36; k++) K[k] = 0.;
    for (int a = 0; a < 8; a++) bf[a] = 0.;
    for (int q = 0; q < 8; q++) {
        double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int a = 0; a < 8; a++)
            for (int d = 0; d < 3; d++)
                for (int e = 0; e < 3; e++) J[d][e] += X[a][d] * g_dN[q * 24 + a * 3 + e];
        const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                           J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        double inv[3][3];
        inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det; inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
        inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det; inv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
        inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
        inv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det; inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
        inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
        double G[8][3];
        for (int a = 0; a < 8; a++)
            for (int d = 0; d < 3; d++) { G[a][d] = 0.; for (int e = 0; e < 3; e++) G[a][d] += inv[e][d] * g_dN[q * 24 + a * 3 + e]; }
        for (int a = 0; a < 8; a++) {
            for (int b = a; b < 8; b++) K[sym_idx(a, b)] += factor * g_w[q] * det * (G[a][0] * G[b][0] + G[a][1] * G[b][1] + G[a][2] * G[b][2]);
            bf[a] += g_Nq[q * 8 + a] * g_w[q] * det;
        }
    }
}
void fill_tables() {
    // same construction as load_q1_tables (isl_engine.cu): C_c[a][b] and sum_q N_a(q) on the 8-point rule
    const isl::Rule R = isl::make_rule(ISL_HEX, 3);
    const isl::Basis B(ISL_HEX, 1);
    std::vector<double> dN(8 * 8 * 3), N(8), Nq(64);
    for (int q = 0; q < 8; q++) { B.eval(&R.p[q * 3], N.data(), &dN[q * 24]); for (int a = 0; a < 8; a++) Nq[q * 8 + a] = N[a]; }
    for (int a = 0; a < 8; a++) { rg_host_nsum[a] = 0.; for (int q = 0; q < 8; q++) rg_host_nsum[a] += Nq[q * 8 + a]; }
    g_w0 = R.w[0];
    for (int k = 0; k < 192; k++) g_dN[k] = dN[k];
    for (int k = 0; k < 64; k++) g_Nq[k] = Nq[k];
    for (int q = 0; q < 8; q++) g_w[q] = R.w[q];
}

template <int A>
void gather_slot_general(const RowMeta& m, const double* sK, int cap, double (&acc)[27], double& body) {
    const int s = m.slot[A];
    if (s != 0xffff) rg_add_slot_general<A>([&](int i) { return sK[i * cap + s]; }, acc, body);
}

}  // namespace

// conn lists the OWNED elements only; the pattern (rowptr, col) may contain more columns (pattern-only halo elements)
extern "C" int emu_rowgather(int64_t n_nodes, int64_t n_elems, const double* coords, const int32_t* conn,
                             const int32_t* node_eqn, const uint8_t* status, const double* presc, const double* values,
                             int64_t n_eqn, const int64_t* rowptr, const int32_t* col, double factor, double f0, int body,
                             int incremental, int store_mode, int rows_per_patch, int NT, int general, double* val, double* rhs,
                             int64_t* stats /* [patches, instances, flagged rows, contiguous blocks, row-by-row blocks] */) {
    fill_tables();
    std::vector<int32_t> heqn((size_t)n_elems * 8), hconn(conn, conn + (size_t)n_elems * 8);
    for (int64_t k = 0; k < n_elems * 8; k++) heqn[k] = node_eqn[conn[k]];
    std::vector<int64_t> hrowptr(rowptr, rowptr + n_eqn + 1);
    std::vector<double> rowxyz((size_t)n_eqn * 3, 0.);
    std::vector<int32_t> perm;
    for (int64_t nd = 0; nd < n_nodes; nd++) {
        const int32_t r = node_eqn[nd];
        if (r < 0) continue;
        for (int d = 0; d < 3; d++) rowxyz[(size_t)r * 3 + d] = coords[(size_t)nd * 3 + d];
        perm.push_back(r);
    }
    const int64_t n_leaves = std::max<int64_t>(1, ((int64_t)perm.size() + rows_per_patch - 1) / rows_per_patch);
    std::vector<int64_t> bounds(n_leaves + 1, 0);
    bounds[n_leaves] = (int64_t)perm.size();
    if (!perm.empty()) rcb_split(perm.data(), rowxyz.data(), 0, (int64_t)perm.size(), (int)n_leaves, 0, bounds.data(), 0);
    PatchHost P;
    P.want_slots = true;
    form_patches(perm, bounds, heqn, hconn, hrowptr, n_eqn, n_nodes, 1 << 30, 1 << 30, P);
    if (!P.lattice) return 1;
    const int n_patches = (int)P.inst_off.size() - 1;
    const size_t nr = P.rows.size();
    // k_row_meta, pass 0 and 1
    std::vector<RowMeta> meta(nr);
    std::vector<int32_t> lift_nodes;
    int counter = 0;
    for (int pid = 0; pid < n_patches; pid++) {
        const int r0 = P.row_off[pid], nrows = P.row_off[pid + 1] - r0;
        const int32_t* ie = P.inst_elem.data() + P.inst_off[pid];
        for (int r = 0; r < nrows; r++) {
            const int32_t g = P.rows[r0 + r];
            RowMeta m; int32_t nbn[27]; bool cnb = false;
            if (!rg_row_tables(g, P.rslot.data() + (size_t)(r0 + r) * 8, ie, conn, node_eqn, status, rowptr, col, m, nbn, cnb)) return 1;
            if (cnb) { m.lift = counter++; lift_nodes.insert(lift_nodes.end(), nbn, nbn + 27); }
            meta[r0 + r] = m;
        }
    }
    stats[0] = n_patches; stats[1] = (int64_t)P.inst_elem.size(); stats[2] = counter; stats[3] = stats[4] = 0;
    const int inst_cap = (P.max_inst + 2) & ~1;
    std::vector<double> sD((size_t)(general == 1 ? 44 : 7) * inst_cap), sX((size_t)P.max_nodes * 3), stage((size_t)(NT / 32) * RG_STAGE);
    for (int pid = 0; pid < n_patches; pid++) {
        const int r0 = P.row_off[pid], nrows = P.row_off[pid + 1] - r0;
        const int n0 = P.node_off[pid], nnodes = P.node_off[pid + 1] - n0;
        const int e0 = P.inst_off[pid], ninst = P.inst_off[pid + 1] - e0;
        for (int n = 0; n < nnodes; n++) for (int d = 0; d < 3; d++) sX[n * 3 + d] = coords[(size_t)P.nodes[n0 + n] * 3 + d];
        // phase 1
        for (int i = 0; i < ninst; i++) {
            const uint16_t* ln = P.lnode.data() + (size_t)(e0 + i) * 8;
            if (general == 1) {  // k_q1hex_rows_general
                double X[8][3], K[36], bf[8];
                for (int a = 0; a < 8; a++) for (int d = 0; d < 3; d++) X[a][d] = sX[ln[a] * 3 + d];
                host_K(X, factor, K, bf);
                for (int k = 0; k < 36; k++) sD[k * inst_cap + i] = K[k];
                for (int a = 0; a < 8; a++) sD[(36 + a) * inst_cap + i] = bf[a];
                continue;
            }
            double D[6], dw;
            rg_instance(&sX[ln[0] * 3], &sX[ln[1] * 3], &sX[ln[3] * 3], &sX[ln[4] * 3], factor, g_w0, D, dw);
            for (int c = 0; c < 6; c++) sD[c * inst_cap + i] = D[c];
            sD[6 * inst_cap + i] = dw;
        }
        if (general != 1) for (int c = 0; c < 7; c++) sD[c * inst_cap + inst_cap - 1] = 0.;   // the zero element
        // phase 2, warp by warp (32 lanes in lock step)
        for (int rb = 0; rb < nrows; rb += NT) {
            for (int warp = 0; warp < NT / 32; warp++) {
                double* st = stage.data() + (size_t)warp * RG_STAGE;
                RowMeta m[32]; int64_t rs[32]; double acc[32][27]; int myn[32]; bool act[32];
                for (int lane = 0; lane < 32; lane++) {
                    const int tid = warp * 32 + lane, r = rb + tid;
                    act[lane] = r < nrows; rs[lane] = 0; myn[lane] = 0;
                    for (int k = 0; k < 27; k++) acc[lane][k] = 0.;
                    if (!act[lane]) continue;
                    m[lane] = meta[r0 + r]; rs[lane] = m[lane].rowstart; myn[lane] = m[lane].nnz & 0x7f;
                    if (m[lane].grow != P.rows[r0 + r] || m[lane].rowstart != rowptr[m[lane].grow]) return 2;
                    double bsum = 0.;
                    if (general == 1) {
                        gather_slot_general<0>(m[lane], sD.data(), inst_cap, acc[lane], bsum); gather_slot_general<1>(m[lane], sD.data(), inst_cap, acc[lane], bsum);
                        gather_slot_general<2>(m[lane], sD.data(), inst_cap, acc[lane], bsum); gather_slot_general<3>(m[lane], sD.data(), inst_cap, acc[lane], bsum);
                        gather_slot_general<4>(m[lane], sD.data(), inst_cap, acc[lane], bsum); gather_slot_general<5>(m[lane], sD.data(), inst_cap, acc[lane], bsum);
                        gather_slot_general<6>(m[lane], sD.data(), inst_cap, acc[lane], bsum); gather_slot_general<7>(m[lane], sD.data(), inst_cap, acc[lane], bsum);
                    } else {
                        int sl[8];
                        for (int a = 0; a < 8; a++) sl[a] = std::min((int)m[lane].slot[a], inst_cap - 1);
                        rg_row_affine([&](int a, int c) { return sD[(size_t)c * inst_cap + sl[a]]; }, acc[lane], bsum);
                    }
                    double lift = 0.;
                    if (m[lane].lift >= 0) {
                        const int32_t* ln = lift_nodes.data() + (size_t)m[lane].lift * 27;
                        for (int k = 0; k < 27; k++) {
                            const int32_t nd = ln[k];
                            if (nd >= 0 && m[lane].pos[k] == 0xff && status[nd] == ISL_CONSTRAINED) {
                                const double gv = incremental ? presc[nd] - values[nd] : presc[nd];
                                lift = std::fma(gv, acc[lane][k], lift);
                            }
                        }
                    }
                    const double v = (body ? f0 * bsum : 0.) - lift;
                    if (v != 0.) rhs[m[lane].grow] += v;
                }
                // rg_write_rows: segments of consecutive rows, staged with the 16-byte phase of their global address, one
                // bulk copy per segment (an odd first / last element leaves as a plain store)
                RgLane L[32]; int total = 0;
                rg_segments_host(myn, rs, L, total);
                std::vector<double> poison(RG_STAGE, std::nan(""));
                std::copy(poison.begin(), poison.end(), st);
                for (int lane = 0; lane < 32; lane++) {
                    if (myn[lane] == 0) continue;
                    const int so = L[lane].off + 2 * L[lane].seg + L[lane].fix;
                    if (so + myn[lane] > RG_STAGE) return 3;
                    if (m[lane].nnz & 0x80) for (int k = 0; k < myn[lane]; k++) st[so + k] = 0.;
                    for (int k = 0; k < 27; k++) if (m[lane].pos[k] != 0xff) st[so + m[lane].pos[k]] = acc[lane][k];
                }
                for (int lane = 0; lane < 32; lane++) {
                    if (!L[lane].head) continue;
                    int64_t g0 = rs[lane];
                    int s0 = L[lane].off + 2 * L[lane].seg + L[lane].fix;
                    int len = (L[lane].next_head < 32 ? L[L[lane].next_head].off : total) - L[lane].off;
                    auto put = [&](int64_t g, int si) { if (store_mode) val[g] = st[si]; else val[g] += st[si]; };
                    if (g0 & 1) { put(g0, s0); g0++; s0++; len--; }
                    if (len & 1) { put(g0 + len - 1, s0 + len - 1); len--; }
                    if (len > 0) {
                        if ((g0 & 1) || (s0 & 1) || (len & 1)) return 4;   // the bulk copy needs 16-byte aligned ends
                        for (int q = 0; q < len; q++) put(g0 + q, s0 + q);
                        stats[3]++;
                    }
                    stats[4]++;
                }
            }
        }
    }
    return 0;
}
