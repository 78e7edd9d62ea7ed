"""Full-size parity for the benchmark workload (BASELINE config 2: structured n^3 Q1-hex Laplace, Dirichlet data from the
Laplace fundamental solution, constant body force) WITHOUT assembling it on the CPU.

On the structured unit-cube mesh with n = 2^p every element is the same cube of edge h = 2^-p, node coordinates and all
intermediate quantities of the reference algorithm scale by exact powers of two, and every pair of ACTIVE neighbours
shares all elements around their common edge/face/cell.  Hence the whole matrix is ONE 27-point stencil,
A[i, i+off] = S[off], and S at mesh size n equals the oracle's stencil at a small mesh size n0 times n0/n (exactly, up to the order in
which the element contributions of a row are summed: a few ulp).
The right-hand side is  f h^3  minus the Dirichlet lift  sum_off S[off] g(node i+off)  over boundary neighbours.
So the oracle, run at n0 = 8, pins every one of the 4.5e8 entries of the 256^3 system."""
import numpy as np

from tests import flows


def oracle_stencil(n0=8, kappa=1.0, quad_deg=3):
    """27-point stencil S0[o], o = (di+1) + 3 (dj+1) + 9 (dk+1), and the body-force entry of an interior row, from the
    oracle on the structured n0^3 mesh; also checks that the oracle's own matrix is that stencil everywhere."""
    from insilico_b200 import engine as E
    c = flows.build_case("laplace_q1_hex", n0, perturb=False)
    c.ops = [("matrix", E.K_LAPLACE, [kappa], quad_deg, 0, 0, True)]
    rp, col, val, _ = c.run_oracle()
    c.ops = [("body", [1.0], quad_deg, 0)]
    body = c.run_oracle()[3]
    m = n0 - 1
    centre = (m // 2) * (1 + m + m * m)
    S = np.zeros(27)
    for p in range(rp[centre], rp[centre + 1]):
        d = int(col[p]) - centre
        dk = int(np.round(d / (m * m))); d -= dk * m * m
        dj = int(np.round(d / m)); di = d - dj * m
        S[(di + 1) + 3 * (dj + 1) + 9 * (dk + 1)] = val[p]
    assert rp[centre + 1] - rp[centre] == 27 and S[13] > 0  # (face neighbours of the trilinear Laplacian are ~0)
    assert np.abs(body - body[centre]).max() <= 1e-14 * abs(body[centre])  # equal up to the summation order
    return S, float(body[centre]), n0


def check_structured_system(n, rowptr, col, val, rhs, dirichlet_fun, S0, body0, n0, tol=1e-12):
    """every entry of the n^3 system against the scaled oracle stencil; returns (max value error, max rhs error),
    both relative"""
    m = n - 1
    n_eqn = m ** 3
    S = S0 * (float(n0) / n)
    body = body0 * (float(n0) / n) ** 3
    assert len(rowptr) == n_eqn + 1 and len(rhs) == n_eqn
    one = 3 - (np.arange(m) == 0).astype(np.int64) - (np.arange(m) == m - 1)
    cnt = (one[:, None, None] * one[None, :, None] * one[None, None, :]).reshape(-1)
    assert np.array_equal(np.diff(rowptr), cnt), "row lengths differ from the 27-point pattern"
    assert rowptr[-1] == (3 * m - 2) ** 3 == len(col) == len(val)
    worst = 0.0
    mm = m * m
    for k in range(m):                                   # one plane of rows at a time bounds the temporaries
        r0, r1 = k * mm, (k + 1) * mm
        p0, p1 = int(rowptr[r0]), int(rowptr[r1])
        rows = np.repeat(np.arange(r0, r1, dtype=np.int64), np.diff(rowptr[r0:r1 + 1]))
        c = col[p0:p1].astype(np.int64)
        di = c % m - rows % m
        dj = (c // m) % m - (rows // m) % m
        dk = c // mm - rows // mm
        assert np.abs(di).max() <= 1 and np.abs(dj).max() <= 1 and np.abs(dk).max() <= 1, "column outside the stencil"
        assert np.all(np.diff(c)[np.diff(rows) == 0] > 0), "columns not ascending"
        o = (di + 1) + 3 * (dj + 1) + 9 * (dk + 1)
        worst = max(worst, float(np.abs(val[p0:p1] - S[o]).max()))
    val_err = worst / np.abs(S).max()
    # right-hand side: body force minus the lift of the boundary values
    ax = np.arange(n + 1) / float(n)
    Z, Y, X = np.meshgrid(ax, ax, ax, indexing="ij")
    G = np.zeros((n + 1,) * 3)
    onb = np.zeros((n + 1,) * 3, dtype=bool)
    onb[0], onb[-1], onb[:, 0], onb[:, -1], onb[:, :, 0], onb[:, :, -1] = True, True, True, True, True, True
    G[onb] = dirichlet_fun(np.stack([X[onb], Y[onb], Z[onb]], axis=1))
    del X, Y, Z
    lift = np.zeros((m, m, m))
    for dk in (-1, 0, 1):
        for dj in (-1, 0, 1):
            for di in (-1, 0, 1):
                lift += S[(di + 1) + 3 * (dj + 1) + 9 * (dk + 1)] * G[1 + dk:n + dk, 1 + dj:n + dj, 1 + di:n + di]
    expected = body - lift.reshape(-1)
    rhs_err = float(np.abs(rhs - expected).max() / np.abs(expected).max())
    assert val_err <= tol and rhs_err <= tol, (val_err, rhs_err)
    return val_err, rhs_err


# ---- perturbed (non-affine) mesh at full size: properties + oracle on sampled sub-boxes --------------------------------
def interior_eqn(n, i, j, k):
    """equation number of the interior node (i, j, k), 1 <= i,j,k <= n-1, in the x-fastest numbering"""
    m = n - 1
    return (i - 1) + m * ((j - 1) + m * (k - 1))


def check_perturbed_system(n, coords, rowptr, col, val, rhs, kappa=1.0, boxes=6, box=5, seed=3, tol=1e-12):
    """System assembled with u = 1 prescribed on the whole boundary and no body force on a perturbed n^3 Q1-hex mesh
    (node numbering x-fastest).  Checks
      * the 27-point pattern,
      * A 1 = rhs (constants are in the kernel of the Laplacian: row sums incl. the lifted boundary columns vanish),
      * symmetry through x^T A y = y^T A x for random vectors,
      * ORACLE PARITY ON SUB-BOXES: for `boxes` random boxes of box^3 elements the oracle assembles the sub-mesh with
        all box-surface nodes constrained; entries A[i, j] between nodes strictly inside the box only receive
        contributions from elements of the box, so they must equal the full system's entries."""
    import scipy.sparse as sp
    from oracle import oracle as orc
    m = n - 1
    one = 3 - (np.arange(m) == 0).astype(np.int64) - (np.arange(m) == m - 1)
    cnt = (one[:, None, None] * one[None, :, None] * one[None, None, :]).reshape(-1)
    assert np.array_equal(np.diff(rowptr), cnt), "row lengths differ from the 27-point pattern"
    A = sp.csr_matrix((val, col.astype(np.int64), rowptr), shape=(m ** 3, m ** 3))
    scale = float(np.abs(val).max())
    assert float(np.abs(A @ np.ones(m ** 3) - rhs).max()) <= 10 * tol * scale
    rng = np.random.default_rng(seed)
    x, y = rng.standard_normal(m ** 3), rng.standard_normal(m ** 3)
    assert abs(x @ (A @ y) - y @ (A @ x)) <= tol * scale * m ** 3
    assert A.diagonal().min() > 0
    worst = 0.0
    X = coords.reshape(n + 1, n + 1, n + 1, 3)           # [k, j, i]
    for _ in range(boxes):
        o = rng.integers(0, n - box + 1, size=3)          # element offset of the box (i0, j0, k0)
        nb = box + 1
        sub = X[o[2]:o[2] + nb, o[1]:o[1] + nb, o[0]:o[0] + nb].reshape(-1, 3).copy()
        ii, jj, kk = np.meshgrid(np.arange(box), np.arange(box), np.arange(box), indexing="ij")  # i fastest below
        e_i, e_j, e_k = ii.transpose(2, 1, 0).reshape(-1), jj.transpose(2, 1, 0).reshape(-1), kk.transpose(2, 1, 0).reshape(-1)
        nid = lambda a, b, c: a + nb * (b + nb * c)
        # hierarchic vertex order of the Q1 hex (HierarchicOrder<HEX,1>: lexicographic {0,1,3,2,4,5,7,6})
        conn = np.stack([nid(e_i, e_j, e_k), nid(e_i + 1, e_j, e_k), nid(e_i + 1, e_j + 1, e_k), nid(e_i, e_j + 1, e_k),
                         nid(e_i, e_j, e_k + 1), nid(e_i + 1, e_j, e_k + 1), nid(e_i + 1, e_j + 1, e_k + 1),
                         nid(e_i, e_j + 1, e_k + 1)], axis=1).astype(np.int64)
        li, lj, lk = np.arange(nb ** 3) % nb, (np.arange(nb ** 3) // nb) % nb, np.arange(nb ** 3) // (nb * nb)
        inner = (li > 0) & (li < box) & (lj > 0) & (lj < box) & (lk > 0) & (lk < box)
        status = (~inner).astype(np.uint8)[:, None]
        eqn, n_sub = orc.number_dofs(status)
        prob = orc.Problem(orc.HEX, 1, sub, conn)
        zeros = np.zeros((nb ** 3, 1))
        prob.set_field(0, 1, 1, nb ** 3, conn, eqn, status, zeros, zeros)
        s = orc.System(n_sub)
        s.stiffness(prob, orc.K_LAPLACE, [kappa], 3, 0, 0, incremental=True, nthreads=1)
        rp_s, col_s, val_s, _ = s.finish()
        # global equation numbers of the box-interior nodes, in sub-mesh numbering order
        gi, gj, gk = li[inner] + o[0], lj[inner] + o[1], lk[inner] + o[2]
        assert gi.min() >= 1 and gi.max() <= n - 1 and gk.min() >= 1 and gk.max() <= n - 1
        g = interior_eqn(n, gi, gj, gk)
        rows_s = np.repeat(np.arange(n_sub), np.diff(rp_s))
        full = np.asarray(A[g[rows_s], g[col_s]]).reshape(-1)
        worst = max(worst, float(np.abs(full - val_s).max()))
        assert len(val_s) > 0
    assert worst <= tol * scale, worst / scale
    return worst / scale
