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

// patches of the leaves [leaf_lo, leaf_hi) appended to P (offsets relative to P); no global scratch arrays, so several
// ranges can be formed concurrently: candidate elements are deduplicated by sort + unique, nodes get their local index
// (order of first appearance over the sorted elements) through a small open-addressing table, rows by binary search
// in the patch's sorted row list
inline void form_patch_range(int64_t leaf_lo, int64_t leaf_hi, const std::vector<int32_t>& row_perm,
                             const std::vector<int64_t>& leaf_bounds, const std::vector<int32_t>& eqn,
                             const std::vector<int32_t>& conn, const std::vector<int64_t>& rowptr,
                             const std::vector<int64_t>& adj_ptr, const std::vector<int32_t>& adj, PatchHost& P) {
    std::vector<uint8_t> amask;
    std::vector<int32_t> hkey, hval;
    for (int64_t leaf = leaf_lo; leaf < leaf_hi; leaf++) {
        // rows of this patch
        const size_t rbase = P.rows.size();
        int entries = 0, nrows = 0;
        for (int64_t k = leaf_bounds[leaf]; k < leaf_bounds[leaf + 1]; k++) {
            const int32_t g = row_perm[k];
            const int nnz = (int)(rowptr[g + 1] - rowptr[g]);
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
            P.inst_elem.insert(P.inst_elem.end(), adj.begin() + adj_ptr[g], adj.begin() + adj_ptr[g + 1]);
        }
        std::sort(P.inst_elem.begin() + ibase, P.inst_elem.end());
        P.inst_elem.erase(std::unique(P.inst_elem.begin() + ibase, P.inst_elem.end()), P.inst_elem.end());
        const size_t ninst = P.inst_elem.size() - ibase;
        size_t hcap = 64;
        while (hcap < ninst * 16) hcap <<= 1;  // <= 8 ninst distinct nodes: load factor <= 1/2
        hkey.assign(hcap, -1); hval.resize(hcap);
        const int32_t* prow = P.rows.data() + rbase;
        int nnodes = 0;
        for (size_t ii = ibase; ii < P.inst_elem.size(); ii++) {
            const int32_t e = P.inst_elem[ii];
            // two local nodes with the same equation (tied / periodic numbering): the phase scatter and the row gather
            // assume distinct columns per local row -> generic (atomic) kernels
            for (int a = 1; a < 8; a++)
                for (int b = 0; b < a; b++)
                    if (eqn[(size_t)e * 8 + a] >= 0 && eqn[(size_t)e * 8 + a] == eqn[(size_t)e * 8 + b]) P.lattice = false;
            for (int a = 0; a < 8; a++) {
                const int32_t nd = conn[(size_t)e * 8 + a];
                size_t hq = ((uint32_t)nd * 2654435761u) & (hcap - 1);
                while (hkey[hq] != -1 && hkey[hq] != nd) hq = (hq + 1) & (hcap - 1);
                if (hkey[hq] == -1) { hkey[hq] = nd; hval[hq] = nnodes++; P.nodes.push_back(nd); }
                P.lnode.push_back((uint16_t)hval[hq]);
                const int32_t ge = eqn[(size_t)e * 8 + a];
                const int32_t* it = ge >= 0 ? std::lower_bound(prow, prow + nrows, ge) : prow + nrows;
                if (it != prow + nrows && *it == ge) {
                    const int l = (int)(it - prow);
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
        P.max_inst = std::max(P.max_inst, (int)ninst);
    }
}

inline void form_patches(const std::vector<int32_t>& row_perm, const std::vector<int64_t>& leaf_bounds,
                         const std::vector<int32_t>& eqn /* [n][8] */,
                         const std::vector<int32_t>& conn /* [n][8] */, const std::vector<int64_t>& rowptr, int64_t n_eqn,
                         int64_t n_nodes, int cap_entries, int cap_nodes, PatchHost& P) {
    const int64_t n = (int64_t)eqn.size() / 8;
#ifdef ISL_PREP_TIMING
    auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tt0 = tnow();
#endif
    // row -> incident element lists; every host thread scans all elements and keeps the rows of its own range
    const int64_t n_leaves = (int64_t)leaf_bounds.size() - 1;
    const int T = (int)std::min<int64_t>(std::max(1u, std::min(16u, std::thread::hardware_concurrency())), std::max<int64_t>(1, n_leaves / 64));
    auto run_threads = [&](auto&& fn) {
        std::vector<std::thread> th;
        for (int t = 0; t + 1 < T; t++) th.emplace_back([&fn, t] { fn(t); });
        fn(T - 1);
        for (auto& x : th) x.join();
    };
    std::vector<int64_t> adj_ptr(n_eqn + 1, 0);
    run_threads([&](int t) {
        const int64_t lo = n_eqn * t / T, hi = n_eqn * (t + 1) / T;
        for (int64_t k = 0; k < n * 8; k++) { const int64_t g = eqn[k]; if (g >= lo && g < hi) adj_ptr[g + 1]++; }
    });
    for (int64_t r = 0; r < n_eqn; r++) adj_ptr[r + 1] += adj_ptr[r];
    std::vector<int32_t> adj(adj_ptr[n_eqn]);
    {
        std::vector<int64_t> cur(adj_ptr.begin(), adj_ptr.end() - 1);
        run_threads([&](int t) {
            const int64_t lo = n_eqn * t / T, hi = n_eqn * (t + 1) / T;
            for (int64_t k = 0; k < n * 8; k++) { const int64_t g = eqn[k]; if (g >= lo && g < hi) adj[cur[g]++] = (int32_t)(k >> 3); }
        });
    }
#ifdef ISL_PREP_TIMING
    const double tt1 = tnow();
#endif
    // contiguous ranges of leaves on the host threads, concatenated in leaf order
    std::vector<PatchHost> part(T);
    run_threads([&](int t) {
        part[t].want_slots = P.want_slots;
        const int64_t lo = n_leaves * t / T, hi = n_leaves * (t + 1) / T;
        const size_t nr = (size_t)(leaf_bounds[hi] - leaf_bounds[lo]), ni = nr * 2 + 1024;  // ~1.5 instances per row
        PatchHost& Q = part[t];
        Q.rows.reserve(nr); Q.soff.reserve(nr + (size_t)(hi - lo)); Q.inst_elem.reserve(ni); Q.nodes.reserve(ni * 2);
        Q.lnode.reserve(ni * 8); Q.lrow.reserve(ni * 8);
        if (Q.want_slots) Q.rslot.reserve(nr * 8);
        form_patch_range(lo, hi, row_perm, leaf_bounds, eqn, conn, rowptr, adj_ptr, adj, Q);
    });
#ifdef ISL_PREP_TIMING
    const double tt2 = tnow();
#endif
    {
        // offsets of every part in the concatenation, then the large arrays are copied by the threads
        std::vector<size_t> bi(T + 1, P.inst_elem.size()), br(T + 1, P.rows.size()), bn(T + 1, P.nodes.size()), bu(T + 1, P.run_start.size()),
            bs(T + 1, P.soff.size()), bq(T + 1, P.run_soff.size());
        for (int t = 0; t < T; t++) {
            const PatchHost& Q = part[t];
            bi[t + 1] = bi[t] + Q.inst_elem.size(); br[t + 1] = br[t] + Q.rows.size(); bn[t + 1] = bn[t] + Q.nodes.size();
            bu[t + 1] = bu[t] + Q.run_start.size(); bs[t + 1] = bs[t] + Q.soff.size(); bq[t + 1] = bq[t] + Q.run_soff.size();
            for (size_t k = 1; k < Q.inst_off.size(); k++) {
                P.inst_off.push_back(Q.inst_off[k] + (int32_t)bi[t]); P.row_off.push_back(Q.row_off[k] + (int32_t)br[t]);
                P.node_off.push_back(Q.node_off[k] + (int32_t)bn[t]); P.run_off.push_back(Q.run_off[k] + (int32_t)bu[t]);
            }
            P.max_entries = std::max(P.max_entries, Q.max_entries); P.max_rows = std::max(P.max_rows, Q.max_rows);
            P.max_nodes = std::max(P.max_nodes, Q.max_nodes); P.max_inst = std::max(P.max_inst, Q.max_inst);
            P.lattice = P.lattice && Q.lattice;
        }
        P.inst_elem.resize(bi[T]); P.rows.resize(br[T]); P.nodes.resize(bn[T]); P.run_start.resize(bu[T]); P.soff.resize(bs[T]);
        P.run_soff.resize(bq[T]); P.lnode.resize(bi[T] * 8); P.lrow.resize(bi[T] * 8);
        if (P.want_slots) P.rslot.resize(br[T] * 8);
        run_threads([&](int t) {
            PatchHost& Q = part[t];
            std::copy(Q.inst_elem.begin(), Q.inst_elem.end(), P.inst_elem.begin() + bi[t]);
            std::copy(Q.rows.begin(), Q.rows.end(), P.rows.begin() + br[t]);
            std::copy(Q.nodes.begin(), Q.nodes.end(), P.nodes.begin() + bn[t]);
            std::copy(Q.run_start.begin(), Q.run_start.end(), P.run_start.begin() + bu[t]);
            std::copy(Q.soff.begin(), Q.soff.end(), P.soff.begin() + bs[t]);
            std::copy(Q.run_soff.begin(), Q.run_soff.end(), P.run_soff.begin() + bq[t]);
            std::copy(Q.lnode.begin(), Q.lnode.end(), P.lnode.begin() + bi[t] * 8);
            std::copy(Q.lrow.begin(), Q.lrow.end(), P.lrow.begin() + bi[t] * 8);
            if (P.want_slots) std::copy(Q.rslot.begin(), Q.rslot.end(), P.rslot.begin() + br[t] * 8);
            Q = PatchHost();
        });
    }
#ifdef ISL_PREP_TIMING
    fprintf(stderr, "[prep] adjacency %.3f s, patches (%d threads) %.3f s, merge %.3f s\n", tt1 - tt0, T, tt2 - tt1, tnow() - tt2);
#endif
    (void)cap_nodes; (void)cap_entries; (void)n_nodes;
}
