"""CPU-only tests: the C ABI library loads and exports what include/insilico_b200.h declares, and the product's
host-side logic (tables, DoF numbering, boundary, mesh generation) agrees with the oracle / reference goldens."""
import os
import re

import numpy as np
import pytest

from insilico_b200 import engine as E
from insilico_b200 import meshgen
from oracle import oracle as orc
from tests import flows
from tests import helpers as H


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(H.ROOT, "include", "insilico_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(isl_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(E.EXPORTED)
    L = E.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.isl_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(E.EngineError, match="no CUDA device"):
        E.Engine(0)


@pytest.mark.parametrize("shape,deg", [(E.HEX, 2), (E.HEX, 3), (E.HEX, 4), (E.QUAD, 3), (E.QUAD, 5), (E.TET, 1),
                                       (E.TET, 2), (E.TET, 3), (E.TET, 4), (E.TET, 5), (E.TRI, 1), (E.TRI, 2),
                                       (E.TRI, 3), (E.TRI, 4), (E.TRI, 5)])
def test_quadrature_tables_equal_oracle(shape, deg):
    w, p = E.quadrature(shape, deg)
    w0, p0 = orc.quadrature(shape, deg)
    assert np.array_equal(w, w0) and np.array_equal(p, p0)


@pytest.mark.parametrize("shape,deg", [(E.HEX, 1), (E.HEX, 2), (E.QUAD, 1), (E.QUAD, 2), (E.QUAD, 3), (E.TET, 1),
                                       (E.TET, 2), (E.TRI, 1), (E.TRI, 2)])
def test_shape_functions_equal_oracle(shape, deg):
    rng = np.random.default_rng(1)
    assert np.array_equal(E.support_points(shape, deg), orc.support_points(shape, deg))
    for _ in range(5):
        xi = rng.random(E.SHAPE_DIM[shape]) / 3
        f, g = E.shape_eval(shape, deg, xi)
        f0, g0 = orc.shape_eval(shape, deg, xi)
        assert np.allclose(f, f0, rtol=0, atol=4e-16) and np.allclose(g, g0, rtol=0, atol=4e-15)


def _mesh(shape, n, permute):
    coords, conn = flows.make_mesh(shape, n, perturb=False, permute=permute)
    return coords, conn


@pytest.mark.parametrize("shape,n,deg,permute", [(E.QUAD, 7, 1, False), (E.QUAD, 7, 2, True), (E.QUAD, 5, 3, True),
                                                 (E.HEX, 4, 1, False), (E.HEX, 4, 2, True), (E.HEX, 3, 3, True),
                                                 (E.TET, 3, 1, True), (E.TET, 3, 2, True), (E.TRI, 6, 2, True)])
def test_dof_numbering_equals_oracle(shape, n, deg, permute):
    coords, conn = _mesh(shape, n, permute)
    ed, nobj = E.dof_generate(shape, 1, conn, deg)
    prob = orc.Problem(shape, 1, coords, conn.astype(np.int64))
    ed0, nobj0 = prob.dof_generate(deg)
    assert nobj == nobj0
    assert np.array_equal(ed.astype(np.int64), ed0)


@pytest.mark.parametrize("deg,nnz", [(1, 3721), (2, 25921), (3, 90601)])
def test_sparsity_goldens_with_product_numbering(deg, nnz):
    shape, coords, conn = H.read_smf(os.path.join(H.REF, "square_20.smf"))
    ed, nobj = E.dof_generate(shape, 1, conn.astype(np.int32), deg)
    pairs = orc.sparsity_pattern(ed.astype(np.int64), nobj)
    gold = H.read_pairs(os.path.join(H.REF, "sparsity.%d.ref.dat" % deg + (".gz" if deg == 3 else "")))
    assert len(pairs) == nnz and np.array_equal(pairs, gold)


@pytest.mark.parametrize("shape,n,deg", [(E.HEX, 3, 1), (E.HEX, 3, 2), (E.TET, 3, 2), (E.QUAD, 5, 2), (E.TRI, 4, 1)])
def test_boundary_and_constraints_equal_oracle(shape, n, deg):
    coords, conn = flows.make_mesh(shape, n, perturb=True, permute=True)
    prob = orc.Problem(shape, 1, coords, conn.astype(np.int64))
    pairs0 = prob.mesh_boundary()
    pairs = E.mesh_boundary(shape, 1, conn)
    assert np.array_equal(pairs, pairs0)
    ed, nobj = E.dof_generate(shape, 1, conn, deg)
    obj, x = E.boundary_dofs(shape, 1, coords, conn, deg, ed, pairs)
    elem0, loc0, x0 = prob.boundary_dof_points(deg, pairs0)
    assert np.array_equal(obj, ed[elem0, loc0])
    assert np.allclose(x, x0, rtol=0, atol=1e-15)
    fun = lambda xx: np.sin(xx.sum(axis=1, keepdims=True))
    st, pr = E.constrain_boundary(shape, 1, coords, conn, deg, 1, ed, nobj, fun)
    st0, pr0 = H.constrain_boundary(prob, deg, 1, ed.astype(np.int64), nobj, fun)
    assert np.array_equal(st, st0) and np.allclose(pr, pr0, rtol=0, atol=1e-15)
    eqn, cnt = E.number_dofs_consecutively(st, init=7)
    eqn0, cnt0 = orc.number_dofs(st0, init=7)
    assert cnt == cnt0 and np.array_equal(eqn, eqn0)


def test_meshgen_equals_reference_recipe():
    c, n, _ = meshgen.unit_cube_hex(3, 4, 5)
    c0, n0 = orc.unit_cube(3, False, 1, 3, 4, 5)
    assert np.array_equal(n, n0) and np.array_equal(c, c0)
    c, n = meshgen.unit_cube_tet(3, 2, 4)
    c0, n0 = orc.unit_cube(3, True, 1, 3, 2, 4)
    assert np.array_equal(n, n0) and np.array_equal(c, c0)
    c, n = meshgen.unit_square_quad(4, 3)
    c0, n0 = orc.unit_cube(2, False, 1, 4, 3)
    assert np.array_equal(n, n0) and np.array_equal(c, c0)
    c, n = meshgen.unit_square_tri(4, 3)
    c0, n0 = orc.unit_cube(2, True, 1, 4, 3)
    assert np.array_equal(n, n0) and np.array_equal(c, c0)
    # slab with global ids is a window of the full mesh
    cs, ns, off = meshgen.unit_cube_hex(3, 4, 6, k0=2, k1=4, global_ids=True)
    cf, nf, _ = meshgen.unit_cube_hex(3, 4, 6)
    assert np.array_equal(ns, nf[2 * 12:4 * 12]) and np.array_equal(cs, cf[off:off + len(cs)])
    # all elements keep a positive Jacobian after the perturbation
    cp = meshgen.perturb_interior(cf, 1.0 / 6, max_dist=0.15)
    assert np.array_equal(cp[meshgen.boundary_node_mask(cf)], cf[meshgen.boundary_node_mask(cf)])
    assert not np.array_equal(cp, cf)


@pytest.mark.parametrize("name", ["laplace_q1_hex", "stvenant_q1_hex", "stokes_p2p1_tet"])
def test_oracle_prestructured_openmp_equals_dynamic(name):
    """registerFields + OpenMP atomic adds (TripletContainer pre-structured mode) gives the dynamic-mode result."""
    c = flows.build_case(name, n=3)
    a = c.run_oracle(register=False, nthreads=1)
    b = c.run_oracle(register=True, nthreads=2)
    r = flows.compare(a, b)
    assert r["pattern_equal"] and r["val_diff"] < 1e-13 and r["rhs_diff"] < 1e-13
    with pytest.raises(RuntimeError, match="Multiple threads"):
        c.run_oracle(register=False, nthreads=2)


def test_oracle_symmetry_and_rigid_body_nullspace():
    """properties of the restated integrals: K symmetric; constants in the kernel of the Laplace stiffness."""
    import scipy.sparse as sp
    c = flows.Case(E.HEX, 1, *flows.make_mesh(E.HEX, 3, True))
    c.add_field(1, 1)
    c.ops = [("matrix", E.K_LAPLACE, [1.0], 3, 0, 0, True)]
    rp, col, val, rhs = c.run_oracle()
    A = sp.csr_matrix((val, col, rp))
    assert abs(A - A.T).max() < 1e-15
    assert np.abs(A @ np.ones(A.shape[0])).max() < 1e-14


def test_cpp_facade_example_builds_and_fails_loudly_without_gpu():
    """include/insilico_b200.hpp (reference template API surface) compiles and links; without a GPU the engine
    aborts with the VERIFY_MSG-style message instead of falling back to the CPU."""
    import subprocess
    import torch
    import __graft_entry__ as g
    g.build()
    exe = os.path.join(H.ROOT, "examples", "heat_dirichlet")
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([exe, "4"], capture_output=True, text=True)
    assert out.returncode != 0 and "no CUDA device available" in out.stderr


# ---- numbering at larger sizes against the reference itself (tools/make_ref_numbering.py) -----------------------------
def _digest(a, dtype):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a, dtype=dtype).tobytes()).hexdigest()


@pytest.mark.parametrize("key", ["stvenant_q2_hex_n16", "neohooke_p2_tet_n16", "stokes_p2p1_tet_n12", "laplace_q2_hex_n16"])
def test_numbering_at_size_equals_reference_run(key):
    """DoF ids per element, status and equation numbers of Q2 hexahedra / P2 tetrahedra / the Taylor-Hood pair on meshes
    with 16 (12) elements per direction, perturbed, in permuted element order: the product's restatement of base/dof
    (isl_dof_generate, isl_mesh_boundary, isl_boundary_dofs, isl_number_dofs through flows.build_case) AND the oracle's
    reproduce what the unmodified reference produced (SHA-256 of the arrays, tests/golden/refrun/numbering_digests.json)"""
    import json
    gold = json.load(open(os.path.join(os.path.dirname(H.REF), "refrun", "numbering_digests.json")))[key]
    name, n = key.rsplit("_n", 1)
    c = flows.build_case(name, int(n), gold["perturb"], gold["permute"])
    assert len(c.conn) == gold["n_elems"]
    prob = orc.Problem(c.shape, 1, c.coords, c.conn.astype(np.int64))
    for f, g in zip(c.fields, gold["fields"]):
        assert f["n_obj"] == g["n_obj"] and int((f["eqn"] >= 0).sum()) == g["n_active"]
        assert _digest(f["elem_dof"], np.int32) == g["elem_dof"], "element -> DoF table differs from the reference's"
        assert _digest(f["status"], np.uint8) == g["status"]
        assert _digest(np.where(f["eqn"] >= 0, f["eqn"], -1), np.int64) == g["eqn"], "equation numbers differ from the reference's"
        ed0, nobj0 = prob.dof_generate(f["fe_deg"])
        assert nobj0 == g["n_obj"] and _digest(ed0, np.int32) == g["elem_dof"]


@pytest.mark.parametrize("shape,n,qdeg", [(E.HEX, 3, 3), (E.TET, 2, 4), (E.QUAD, 5, 3), (E.TRI, 4, 4)])
def test_boundary_surface_elements_and_points_equal_oracle(shape, n, qdeg):
    """host side of the surface terms (isl_boundary_surface, isl_surface_points) against the oracle's restatement of
    generateBoundaryMesh / Geometry / SurfaceNormal (pinned on the reference run by the neumann_* fixtures)"""
    from oracle import oracle as orc
    from tests import flows
    coords, conn = flows.make_mesh(shape, n, True, True)
    pairs = E.mesh_boundary(shape, 1, conn)
    ss, de, sx, sp = E.boundary_surface(shape, 1, coords, conn, pairs)
    prob = orc.Problem(shape, 1, coords, conn.astype(np.int64))
    de_o, sx_o, sp_o = prob.boundary_surface(prob.mesh_boundary())
    assert np.array_equal(de, de_o) and np.array_equal(sx, sx_o) and np.array_equal(sp, sp_o)
    x, nr, dg = E.surface_points(ss, 1, sx, qdeg)
    xo, nro, dgo = orc.surface_points(ss, 1, sx, qdeg)
    assert np.abs(x - xo).max() <= 1e-15 and np.abs(nr - nro).max() <= 1e-15 and np.abs(dg - dgo).max() <= 1e-15 * dg.max()
    assert np.allclose(np.linalg.norm(nr, axis=-1), 1.0, atol=1e-14)
    # outward: the normal points away from the centre of the unit square / cube on every boundary face
    assert np.all(((x - 0.5) * nr).sum(axis=-1) > 0)
