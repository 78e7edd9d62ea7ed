"""Oracle parity on sampled sub-meshes for ANY workload at ANY size (test infrastructure).

A matrix entry A[r, c] between two DoF components that only elements of a sub-mesh S touch receives contributions from
S alone, so the oracle can assemble S by itself (all other DoFs of S constrained, the current field values kept) and the
entry must equal the one of the full system assembled by the CUDA engine.  Used at the sizes of BASELINE configs 3 and 4
(tests/test_zy_full_size.py), where the oracle cannot assemble the whole mesh."""
import numpy as np

from oracle import oracle as orc


def _lookup(rowptr, col, val, r, c):
    """values of the entries (r[k], c[k]) of a CSR matrix with ascending columns; NaN where absent"""
    out = np.full(len(r), np.nan)
    for k in range(len(r)):
        a, b = rowptr[r[k]], rowptr[r[k] + 1]
        j = a + np.searchsorted(col[a:b], c[k])
        if j < b and col[j] == c[k]:
            out[k] = val[j]
    return out


def check_subboxes(w, rowptr, col, val, rhs=None, n_boxes=3, half_width=0.12, seed=11, tol=1e-12, matrix_ops=None):
    """w: insilico_b200.workloads.Workload (single field pair per matrix op); (rowptr, col, val) the engine's system.
    n_boxes cubes of half width `half_width` around random centres: S = elements with their centroid inside."""
    rng = np.random.default_rng(seed)
    cen = w.coords[w.conn].mean(axis=1)
    ops = [op for op in w.ops if op[0] == "matrix"] if matrix_ops is None else matrix_ops
    scale = float(np.abs(val).max())
    worst, compared = 0.0, 0
    counts = [np.bincount(f["elem_dof"].reshape(-1), minlength=f["n_obj"]) for f in w.fields]
    for _ in range(n_boxes):
        c0 = rng.uniform(half_width, 1.0 - half_width, size=w.dim)
        sel = np.nonzero(np.all(np.abs(cen - c0) <= half_width, axis=1))[0]
        assert len(sel) > 0
        nodes = np.unique(w.conn[sel])
        g2l = np.full(len(w.coords), -1, dtype=np.int64); g2l[nodes] = np.arange(len(nodes))
        prob = orc.Problem(w.shape, w.geom_deg, w.coords[nodes], g2l[w.conn[sel]])
        n_sub, maps = 0, []
        for i, f in enumerate(w.fields):
            objs = np.unique(f["elem_dof"][sel])
            o2l = np.full(f["n_obj"], -1, dtype=np.int64); o2l[objs] = np.arange(len(objs))
            sub_cnt = np.bincount(f["elem_dof"][sel].reshape(-1), minlength=f["n_obj"])[objs]
            inner_obj = sub_cnt == counts[i][objs]                    # touched by elements of S only
            act = inner_obj[:, None] & (f["status"][objs] == 0)      # and ACTIVE in the full problem
            status = np.where(act, 0, 1).astype(np.uint8)
            eqn, nn = orc.number_dofs(status, init=n_sub)
            # prescribed = current value: the incremental lift of the sub-problem is irrelevant, only the matrix is compared
            prob.set_field(i, f["fe_deg"], f["ds"], len(objs), o2l[f["elem_dof"][sel]], eqn, status, f["values"][objs], f["values"][objs])
            maps.append((objs, eqn, act))
            n_sub += nn
        if n_sub == 0:
            continue
        s = orc.System(n_sub)
        for op in ops:
            s.stiffness(prob, op[1], op[2], op[3], op[4], op[5], incremental=True, nthreads=1)
        rp_s, col_s, val_s, _ = s.finish()
        # local equation -> global equation
        l2g = np.full(n_sub, -1, dtype=np.int64)
        for (objs, eqn, act), f in zip(maps, w.fields):
            l2g[eqn[act]] = f["eqn"][objs][act]
        assert l2g.min() >= 0
        rows_s = np.repeat(np.arange(n_sub), np.diff(rp_s))
        full = _lookup(rowptr, col, val, l2g[rows_s], l2g[col_s])
        assert not np.isnan(full).any(), "an entry of the sub-system is missing in the engine's pattern"
        worst = max(worst, float(np.abs(full - val_s).max()))
        compared += len(val_s)
    assert compared > 0
    assert worst <= tol * scale, (worst / scale, compared)
    return worst / scale, compared
