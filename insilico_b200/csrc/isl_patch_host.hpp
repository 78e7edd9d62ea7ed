// =============================================================================
// isl_patch_host.hpp -- host-side preprocessing of the owner-computes patch kernels (pure C++, no CUDA):
// recursive coordinate bisection of the rows into compact boxes and the element-instance lists of every box.
// Included by isl_patch.cuh (inside the engine's anonymous namespace) and by the host emulation under tests/emu/.
// The including file provides <vector>, <algorithm>, <thread>, <cstdint>.
// =============================================================================
#pragma once

struct PatchHost {
    std::vector<int32_t> inst_off{0}, row_off{0}, node_off{0}, run_off{0}, rows, nodes, inst_elem;
    std::vector<uint32_t> soff, run_soff;
    std::vector<int64_t> run_start;
    std::vector<uint16_t> lnode, lrow;
    std::vector<uint16_t> rslot;  // (want_slots) per owned row: patch-local instance that has the row as local node a
    bool want_slots = false;
    int max_entries = 0, max_rows = 0, max_nodes = 0, max_inst = 0;
    bool lattice = true;
};

// rows in Morton order are cut into patches of <= rows_per_patch rows / cap_entries matrix entries; each patch lists
// every element touching one of its rows (owner computes)
// recursive coordinate bisection of the rows into `leaves` compact boxes of (almost) equal size; idx is permuted in
// place, leaf k covers idx[bounds[k] .. bounds[k+1])
inline void rcb_split(int32_t* idx, const double* xyz, int64_t lo, int64_t hi, int leaves, int64_t leaf0, int64_t* bounds,
                      int depth) {
    if (leaves <= 1) { bounds[leaf0] = lo; std::sort(idx + lo, idx + hi); return; }
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int64_t k = lo; k < hi; k++)
        for (int d = 0; d < 3; d++) { const double v = xyz[(size_t)idx[k] * 3 + d]; mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v); }
    int d = 0;
    if (mx[1] - mn[1] > mx[d] - mn[d]) d = 1;
    if (mx[2] - mn[2] > mx[d] - mn[d]) d = 2;
    const int l1 = leaves / 2;
    const int64_t mid = lo + (hi - lo) * l1 / leaves;
    std::nth_element(idx + lo, idx + mid, idx + hi, [&](int32_t a, int32_t b) {
        const double va = xyz[(size_t)a * 3 + d], vb = xyz[(size_t)b * 3 + d];
        return va < vb || (va == vb && a < b);
    });
    if (depth < 4) {
        std::thread t([=] { rcb_split(idx, xyz, lo, mid, l1, leaf0, bounds, depth + 1); });
        rcb_split(idx, xyz, mid, hi, leaves - l1, leaf0 + l1, bounds, depth + 1);
        t.join();
    } else {
        rcb_split(idx, xyz, lo, mid, l1, leaf0, bounds, depth + 1);
        rcb_split(idx, xyz, mid, hi, leaves - l1, leaf0 + l1, bounds, depth + 1);
    }
}

inline void form_patches(const std::vector<int32_t>& row_perm, const std::vector<int64_t>& leaf_bounds,
                         const std::vector<int32_t>& eqn /* [n][8] */,
                         const std::vector<int32_t>& conn /* [n][8] */, const std::vector<int64_t>& rowptr, int64_t n_eqn,
                         int64_t n_nodes, int cap_entries, int cap_nodes, PatchHost& P) {
    const int64_t n = (int64_t)eqn.size() / 8;
    // row -> incident (element, local index) lists
    std::vector<int64_t> adj_ptr(n_eqn + 1, 0);
    for (int64_t k = 0; k < n * 8; k++) if (eqn[k] >= 0) adj_ptr[eqn[k] + 1]++;
    for (int64_t r = 0; r < n_eqn; r++) adj_ptr[r + 1] += adj_ptr[r];
    std::vector<int32_t> adj(adj_ptr[n_eqn]);
    {
        std::vector<int64_t> cur(adj_ptr.begin(), adj_ptr.end() - 1);
        for (int64_t e = 0; e < n; e++)
            for (int a = 0; a < 8; a++) { const int32_t g = eqn[e * 8 + a]; if (g >= 0) adj[cur[g]++] = (int32_t)e; }
    }
    std::vector<int32_t> row_stamp(n_eqn, -1), row_l(n_eqn, 0), el_stamp(n, -1), node_stamp(n_nodes, -1), node_l(n_nodes, 0);
    std::vector<uint8_t> amask;
    int pid = 0;
    const int64_t n_leaves = (int64_t)leaf_bounds.size() - 1;
    for (int64_t leaf = 0; leaf < n_leaves; leaf++) {
        // rows of this patch
        const size_t rbase = P.rows.size();
        int entries = 0, nrows = 0;
        for (int64_t k = leaf_bounds[leaf]; k < leaf_bounds[leaf + 1]; k++) {
            const int32_t g = row_perm[k];
            const int nnz = (int)(rowptr[g + 1] - rowptr[g]);
            row_stamp[g] = pid; row_l[g] = nrows;
            P.rows.push_back(g); P.soff.push_back((uint32_t)entries);
            entries += nnz; nrows++;
        }
        if (nrows == 0) continue;
        P.soff.push_back((uint32_t)entries);
        {   // runs of consecutive global rows (rows are sorted by global id inside a patch)
            int32_t prev = -2;
            for (int r = 0; r < nrows; r++) {
                const int32_t g = P.rows[rbase + r];
                if (g != prev + 1) { P.run_start.push_back(rowptr[g]); P.run_soff.push_back(P.soff[P.soff.size() - 1 - nrows + r]); }
                prev = g;
            }
            P.run_soff.push_back((uint32_t)entries);
            P.run_off.push_back((int32_t)P.run_start.size());
        }
        amask.assign(nrows, 0);
        if (P.want_slots) P.rslot.resize((rbase + (size_t)nrows) * 8, (uint16_t)0xffff);
        // elements touching those rows (sorted by element id: consecutive lanes then work on neighbouring elements,
        // which keeps their shared-memory accesses on different banks), nodes of those elements
        const size_t ibase = P.inst_elem.size();
        for (int r = 0; r < nrows; r++) {
            const int32_t g = P.rows[rbase + r];
            for (int64_t j = adj_ptr[g]; j < adj_ptr[g + 1]; j++) {
                const int32_t e = adj[j];
                if (el_stamp[e] == pid) continue;
                el_stamp[e] = pid;
                P.inst_elem.push_back(e);
            }
        }
        std::sort(P.inst_elem.begin() + ibase, P.inst_elem.end());
        int nnodes = 0;
        for (size_t ii = ibase; ii < P.inst_elem.size(); ii++) {
            const int32_t e = P.inst_elem[ii];
            for (int a = 0; a < 8; a++) {
                const int32_t nd = conn[(size_t)e * 8 + a];
                if (node_stamp[nd] != pid) { node_stamp[nd] = pid; node_l[nd] = nnodes++; P.nodes.push_back(nd); }
                P.lnode.push_back((uint16_t)node_l[nd]);
                const int32_t ge = eqn[(size_t)e * 8 + a];
                if (ge >= 0 && row_stamp[ge] == pid) {
                    const int l = row_l[ge];
                    P.lrow.push_back((uint16_t)l);
                    if (amask[l] & (1u << a)) P.lattice = false;  // two elements see this row as local row a
                    amask[l] |= (uint8_t)(1u << a);
                    if (P.want_slots) P.rslot[(rbase + (size_t)l) * 8 + a] = (uint16_t)(ii - ibase);
                } else P.lrow.push_back((uint16_t)0xffff);
            }
        }
        P.inst_off.push_back((int32_t)P.inst_elem.size());
        P.row_off.push_back((int32_t)P.rows.size());
        P.node_off.push_back((int32_t)P.nodes.size());
        P.max_entries = std::max(P.max_entries, entries);
        P.max_rows = std::max(P.max_rows, nrows);
        P.max_nodes = std::max(P.max_nodes, nnodes);
        P.max_inst = std::max(P.max_inst, (int)(P.inst_elem.size() - ibase));
        pid++;
    }
    (void)cap_nodes; (void)cap_entries;
}
