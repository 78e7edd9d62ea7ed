"""Pins the CPU oracle against the reference's own golden files.  CPU only."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import oracle as orc
from tests import helpers as H


def test_hierarchic_order_tables():
    # check values derived from base/mesh/HierarchicOrder.hpp
    assert list(orc.hierarchic_order(orc.QUAD, 1)) == [0, 1, 3, 2]
    assert list(orc.hierarchic_order(orc.QUAD, 2)) == [0, 4, 1, 7, 8, 5, 3, 6, 2]
    assert list(orc.hierarchic_order(orc.HEX, 1)) == [0, 1, 3, 2, 4, 5, 7, 6]
    assert list(orc.hierarchic_order(orc.HEX, 2)) == [0, 8, 1, 11, 20, 9, 3, 10, 2, 16, 22, 17, 25, 26, 23, 19, 24,
                                                       18, 4, 12, 5, 15, 21, 13, 7, 14, 6]


def test_quadrature_point_counts_and_weights():
    for shape, deg, n in [(orc.HEX, 3, 8), (orc.HEX, 4, 27), (orc.HEX, 5, 27), (orc.TET, 4, 11), (orc.TET, 2, 4),
                          (orc.TET, 3, 5), (orc.QUAD, 3, 4), (orc.TRI, 2, 3)]:
        w, p = orc.quadrature(shape, deg)
        assert len(w) == n
        vol = {orc.HEX: 1.0, orc.QUAD: 1.0, orc.TET: 1.0 / 6.0, orc.TRI: 0.5}[shape]
        assert abs(w.sum() - vol) < 2e-14 * max(1, n) + 1e-6 * (shape in (orc.TET, orc.TRI, orc.HEX))
    # tensor ordering x-fastest, weight (w_x*w_y)*w_z  (quad/TensorProduct.hpp:122-177)
    w, p = orc.quadrature(orc.HEX, 3)
    assert p[0].tolist() == [0.788675134594813] * 3
    assert p[1].tolist() == [0.211324865405187, 0.788675134594813, 0.788675134594813]
    assert w[0] == (0.5 * 0.5) * 0.5


def test_shape_functions_partition_of_unity_and_kronecker():
    rng = np.random.default_rng(0)
    for shape, deg in [(orc.QUAD, 1), (orc.QUAD, 2), (orc.QUAD, 3), (orc.HEX, 1), (orc.HEX, 2), (orc.TET, 1),
                       (orc.TET, 2), (orc.TRI, 1), (orc.TRI, 2)]:
        sp_ = orc.support_points(shape, deg)
        n = len(sp_)
        V = np.array([orc.shape_eval(shape, deg, sp_[i])[0] for i in range(n)])
        assert np.allclose(V, np.eye(n), atol=1e-13), (shape, deg)
        xi = rng.random(orc.SHAPE_DIM[shape]) * 0.3
        f, g = orc.shape_eval(shape, deg, xi)
        assert abs(f.sum() - 1) < 1e-13 and np.abs(g.sum(axis=0)).max() < 1e-12
        # gradient vs central differences
        for d in range(len(xi)):
            e = np.zeros_like(xi); e[d] = 1e-6
            fd = (orc.shape_eval(shape, deg, xi + e)[0] - orc.shape_eval(shape, deg, xi - e)[0]) / 2e-6
            assert np.allclose(fd, g[:, d], atol=1e-7)


def test_face_numbering_worked_example():
    # base/dof/generateDoFIndicesFromFaces.hpp:141-160 : 2x2 quads, Q2 field on Q1 geometry
    coords = np.array([[i % 3 * 0.5, i // 3 * 0.5] for i in range(9)])
    conn = np.array([[0, 1, 4, 3], [1, 2, 5, 4], [3, 4, 7, 6], [4, 5, 8, 7]])
    prob = orc.Problem(orc.QUAD, 1, coords, conn)
    ed, n = prob.dof_generate(2)
    assert n == 25
    assert ed.tolist() == [[0, 1, 2, 3, 9, 10, 11, 12, 21], [1, 4, 5, 2, 13, 14, 15, 10, 22],
                           [3, 2, 6, 7, 11, 16, 17, 18, 23], [2, 5, 8, 6, 15, 19, 20, 16, 24]]


@pytest.mark.parametrize("deg,nnz", [(1, 3721), (2, 25921), (3, 90601)])
def test_sparsity_goldens(deg, nnz):
    # reference/03-doFHandler/sparsity.{1,2,3}.ref.dat from square_20.smf (doFHandler.cpp:86-107)
    shape, coords, conn = H.read_smf(os.path.join(H.REF, "square_20.smf"))
    prob = orc.Problem(shape, 1, coords, conn)
    ed, nobj = prob.dof_generate(deg)
    pairs = orc.sparsity_pattern(ed, nobj)
    n = len(pairs)
    name = "sparsity.%d.ref.dat" % deg + (".gz" if deg == 3 else "")
    gold = H.read_pairs(os.path.join(H.REF, name))
    assert n == nnz == len(gold)
    assert np.array_equal(pairs, gold)


def test_measure_golden():
    # reference/02-areaVolume/measure.ref.dat : P2 tets of a sphere, tet rules of degree 1..5
    shape, coords, conn = H.read_smf(os.path.join(H.REF, "sphere.tetrahedron.smf"))
    assert shape == orc.TET and conn.shape[1] == 10
    prob = orc.Problem(shape, 2, coords, conn)
    gold = [8.90126e-05, 0.000240783, 0.000241114, 0.000241114, 0.000241114]
    for deg, g in zip(range(1, 6), gold):
        v = prob.measure(deg)
        err = abs(4. / 3. * np.pi - v)  # areaVolume.cpp:170-177 prints |exact - computed|, radius 1
        assert float("%.6g" % err) == pytest.approx(g, rel=2e-6), (deg, v)


def _linear_elastic(smf, dim):
    """reference/06-elastic/linearElastic.cpp:83-213 flow on the oracle."""
    E, nu = 1000.0, 0.25
    lam = E * nu / (1. + nu) / (1. - 2. * nu)
    mu = E / 2. / (1. + nu)
    shape, coords, conn = H.read_smf(smf)
    prob = orc.Problem(shape, 1, coords, conn)
    ed, nobj = prob.dof_generate(1)
    y0 = np.full(dim, -0.1)
    direction = np.zeros(dim); direction[1] = 1.0
    sol = lambda x: H.fund_sol_elastostatic(x, y0, direction, lam, mu)
    status, presc = H.constrain_boundary(prob, 1, dim, ed, nobj, sol)
    eqn, ndof = orc.number_dofs(status)
    values = np.zeros((nobj, dim))
    prob.set_field(0, 1, dim, nobj, ed, eqn, status, presc, values)
    sysm = orc.System(ndof)
    sysm.stiffness(prob, orc.K_HYPEL_STVENANT, [lam, mu], 3, 0, 0)
    rowptr, col, val, rhs = sysm.finish()
    A = sp.csr_matrix((val, col, rowptr), shape=(ndof, ndof))
    assert abs(A - A.T).max() < 1e-9 * abs(A).max()
    x = spla.spsolve(A.tocsc(), rhs) if ndof > 0 else np.zeros(0)
    # dof::setDoFsFromSolver (base/dof/Distribute.hpp): active <- solution, constrained <- prescribed
    u = np.where(status == 0, 0.0, presc)
    act = status == 0
    u[act] = x[eqn[act]]
    prob.set_field_values(0, u)
    xq = prob.quad_points_x(0, 3).reshape(-1, dim)
    uref = sol(xq).reshape(prob.n_elems, -1, dim)
    return prob.l2_error(0, 3, uref)


@pytest.mark.parametrize("name,gold", [("cube.002.smf", 1.89676e-05), ("cube.004.smf", 3.33718e-06),
                                       ("cube.008.smf", 7.28627e-07)])
def test_linear_elastic_3d_golden(name, gold):
    err = _linear_elastic(os.path.join(H.REF, name), 3)
    assert float("%.6g" % err) == pytest.approx(gold, rel=2e-6), err


@pytest.mark.parametrize("name,gold", [("quad.002.smf", 1.54522e-05), ("quad.005.smf", 2.55781e-06),
                                       ("quad.010.smf", 6.58286e-07)])
def test_linear_elastic_2d_golden(name, gold):
    err = _linear_elastic(os.path.join(H.REF, name), 2)
    assert float("%.6g" % err) == pytest.approx(gold, rel=2e-6), err


def test_unit_cube_matches_fixture():
    # tools/meshGeneration/unitCube restated in memory reproduces reference/06-elastic/cube.004.smf
    shape, coords, conn = H.read_smf(os.path.join(H.REF, "cube.004.smf"))
    c2, n2 = orc.unit_cube(3, False, 1, 4, 4, 4)
    assert np.array_equal(conn, n2)
    assert np.allclose(coords, c2, atol=1e-6)  # SMF text holds 6 digits
    shape, coords, conn = H.read_smf(os.path.join(H.REF, "square_20.smf"))
    c2, n2 = orc.unit_cube(2, False, 1, 20, 20)
    assert np.array_equal(conn, n2)


def test_compressible_newton_history_golden():
    """reference/06-elastic/compressible.cpp:245-320 (displacement controlled, inputCompRefD.dat) on quad.020.smf
    (identical to square_20.smf): HyperElastic<NeoHookeanCompressible> tangent + residual on Q2 quads, five load steps
    of Newton iterations.  Golden compRefOutD.dat: #dofs 3239 and |F|, |x| per iteration (6 digits)."""
    E_, nu = 1000.0, 0.3
    lam = E_ * nu / (1. + nu) / (1. - 2. * nu)
    mu = E_ / 2. / (1. + nu)
    pull, load_steps, tol, max_iter = 3.0, 5, 1e-12, 30
    shape, coords, conn = H.read_smf(os.path.join(H.REF, "square_20.smf"))
    prob = orc.Problem(shape, 1, coords, conn)
    ed, nobj = prob.dof_generate(2)
    # PulledSheet<2>::dirichletBC (PulledSheet.hpp:33-57): left side fixed, right side pulled in x
    pairs = prob.mesh_boundary()
    elem, loc, x = prob.boundary_dof_points(2, pairs)
    status = np.zeros((nobj, 2), dtype=np.uint8)
    presc = np.zeros((nobj, 2))
    first_pull = pull / load_steps
    for k in range(len(elem)):
        o = ed[elem[k], loc[k]]
        if abs(x[k, 0]) < 1e-6:
            status[o, :] = 1; presc[o, :] = 0.0
        if abs(x[k, 0] - 1.0) < 1e-6 and status[o, 0] == 0:
            status[o, 0] = 1; presc[o, 0] = first_pull
    eqn, ndof = orc.number_dofs(status)
    assert ndof == 3239
    gold = [l.split() for l in open(os.path.join(H.REF, "compRefOutD.dat")) if not l.startswith("#")]
    values = np.zeros((nobj, 2))
    g = 0
    for step in range(load_steps):
        presc *= 1.0 if step == 0 else (step + 1) / step          # dof::scaleConstraints
        for it in range(max_iter):
            prob.set_field(0, 2, 2, nobj, ed, eqn, status, presc, values)
            s = orc.System(ndof)
            s.residual(prob, orc.K_HYPEL_NEOHOOKE, [lam, mu], 3, 0, 0)
            s.stiffness(prob, orc.K_HYPEL_NEOHOOKE, [lam, mu], 3, 0, 0, incremental=True)
            conv1 = s.rhs_norm()
            rowptr, col, val, rhs = s.finish()
            row = gold[g]; g += 1
            assert int(row[0]) == step and int(row[1]) == it
            assert float("%.6g" % conv1) == pytest.approx(float(row[2]), rel=3e-5, abs=1e-13), (step, it, conv1)
            if conv1 < tol * E_:
                break
            A = sp.csr_matrix((val, col, rowptr), shape=(ndof, ndof))
            dx = spla.spsolve(A.tocsc(), rhs)
            act = status == 0
            values[act] += dx[eqn[act]]                              # dof::addToDoFsFromSolver
            values[status == 1] = presc[status == 1]
            conv2 = np.linalg.norm(dx) / ndof
            assert float("%.6g" % conv2) == pytest.approx(float(row[3]), rel=3e-5), (step, it, conv2)
            if conv2 < tol:
                break
    assert g == len(gold)
