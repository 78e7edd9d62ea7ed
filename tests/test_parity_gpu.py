"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle on the same seeded inputs.
Bar: CSR pattern bit-exact; values and rhs within 1e-12 (metric of SURVEY 8d: |a-b| / max(|a|,|b|,max_row|A|))."""
import numpy as np
import pytest
import scipy.sparse as sp

from insilico_b200 import engine as E
from tests import flows
from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def eng():
    e = E.Engine(0)
    yield e
    e.close()


CASES = [
    ("laplace_q1_hex", 8, True, False), ("laplace_q1_hex", 5, False, False), ("laplace_q1_hex", 12, True, True),
    ("laplace_q1_hex_values", 6, True, True), ("laplace_q2_hex", 3, True, False), ("laplace_p1_tet", 5, True, True),
    ("laplace_q1_quad", 9, True, False), ("laplace_p2_tri", 6, True, True), ("vector_laplace_q1_hex", 5, True, False),
    ("stvenant_q1_hex", 4, True, False), ("stvenant_q2_hex", 2, True, False), ("neohooke_q1_hex", 3, True, True),
    ("stvenant_q1_quad", 7, True, False), ("neohooke_p2_tet", 3, True, True), ("stokes_p2p1_tet", 3, True, False),
    ("stokes_q2q1_hex", 2, True, False), ("stokes_q2q1_quad", 5, True, True),
    # base::kernel::Mass + a stiffness matrix into ONE system: the Q1 row kernel then runs in accumulate mode (bulk
    # reduction) on the structured mesh and through the two-kernel general path on the perturbed one
    ("mass_q1_hex", 7, False, False), ("mass_q1_hex", 6, True, True), ("mass_p2_tet_vector", 3, True, True),
    ("stvenant_q2_hex", 3, True, True), ("stvenant_p2_tet", 3, True, True),
    # heat::Laplace with a conductivity function sampled per quadrature point (isl_assemble_matrix_sampled)
    ("laplace_q1_hex_kappafun", 6, True, True), ("laplace_p2_tet_kappafun", 3, True, False),
    # fluid::Convection on the tuple (u, u, u): advection velocity = third field (isl_assemble_matrix_aux / _residual_aux)
    ("convection_q1_hex", 5, True, True), ("convection_q2_quad", 6, True, False), ("convection_q1_hex_at_rest", 3, True, False),
    # surface (Neumann) terms in one launch: isl_assemble_neumann with a sampled f(x, n), a pressure, a constant traction;
    # quadrilateral / triangle / line surface elements, Dirichlet on a part of the boundary, linear constraints
    ("neumann_q1_hex", 6, True, True), ("neumann_q1_hex", 5, False, False), ("neumann_p2_tet_solid", 3, True, True),
    ("neumann_p2_tri", 6, True, False), ("neumann_q1_quad_solid", 7, True, True),
]


@pytest.mark.parametrize("name,n,perturb,permute", CASES)
def test_assembly_parity(eng, name, n, perturb, permute):
    c = flows.build_case(name, n, perturb, permute)
    ref = c.run_oracle()
    out = c.run_engine(eng=eng)
    r = flows.compare(ref, out)
    assert r["pattern_equal"], r
    assert r["val_diff"] <= TOL and r["rhs_diff"] <= TOL, r
    assert r["nnz"] > 0


GENERIC_CASES = [
    ("laplace_q2_hex", 3, True, False), ("laplace_p1_tet", 5, True, True), ("laplace_q1_quad", 9, True, False),
    ("laplace_p2_tri", 6, True, True), ("vector_laplace_q1_hex", 5, True, False), ("stokes_p2p1_tet", 3, True, False),
    ("stokes_p2p1_tet", 4, True, True), ("stokes_q2q1_hex", 2, True, False), ("stokes_q2q1_quad", 5, True, True),
    ("mass_q1_hex", 6, True, True), ("mass_p2_tet_vector", 3, True, True), ("convection_q1_hex", 5, True, True),
    ("convection_q2_quad", 6, True, False),
    # conductivity sampled per quadrature point (isl_assemble_matrix_sampled)
    ("laplace_q1_hex_kappafun", 6, True, True), ("laplace_p2_tet_kappafun", 3, True, False),
]


@pytest.mark.parametrize("gather,tile", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("name,n,perturb,permute", GENERIC_CASES)
def test_generic_kernels_atomic_and_atomic_free_scatter(name, n, perturb, permute, gather, tile):
    """k_tangent with the atomic scatter through the slot / node-block maps (gen_gather 0) and with the element matrices
    stored and the CSR rows gathered by sub-warps (gen_gather 1, isl_gather.cuh), per-entry loops (gen_tile 0) and register
    strips (gen_tile 1, k_tangent_strips): all against the oracle, registered and dynamic pattern; the gathered system is
    reproducible bit for bit"""
    e = E.Engine(0)
    try:
        e.set_option("gen_gather", gather)
        e.set_option("gen_tile", tile)
        c = flows.build_case(name, n, perturb, permute)
        ref = c.run_oracle()
        out = c.run_engine(eng=e)
        r = flows.compare(ref, out)
        assert r["pattern_equal"] and r["val_diff"] <= TOL and r["rhs_diff"] <= TOL and r["nnz"] > 0, r
        again = c.run_engine(eng=e, register=True)
        r = flows.compare(ref, again)
        assert r["pattern_equal"] and r["val_diff"] <= TOL and r["rhs_diff"] <= TOL, r
        if gather:
            assert np.array_equal(out[2], again[2]), "the gathered matrix is not reproducible bit for bit"
    finally:
        e.close()


@pytest.mark.parametrize("name", ["laplace_q1_hex", "stokes_p2p1_tet"])
def test_registered_pattern_equals_dynamic(eng, name):
    """solver.registerFields first (pre-structured) or pattern discovered by the assembly calls: same system."""
    c = flows.build_case(name, 3)
    a = c.run_engine(eng=eng, register=False)
    b = c.run_engine(eng=eng, register=True)
    r = flows.compare(a, b)
    assert r["pattern_equal"] and r["val_diff"] <= TOL and r["rhs_diff"] <= TOL


def test_newton_loop_reuses_pattern_and_matches_oracle(eng):
    """fresh solver per Newton iteration (compressible.cpp:265): pattern cached, values re-zeroed, field updated."""
    c = flows.build_case("neohooke_q1_hex", 3)
    first = c.run_engine(eng=eng)
    launches0 = eng.kernel_launches
    f = c.fields[0]
    f["values"] = f["values"] * 0.5 + 0.001
    ref = c.run_oracle()
    eng.update_field(0, values=np.ascontiguousarray(f["values"]))
    eng.new_solver(c.n_eqn)
    for op in c.ops:
        if op[0] == "matrix":
            eng.stiffness_matrix_computation(op[1], op[2], op[3], op[4], op[5], incremental=op[6])
        else:
            eng.compute_residual_forces(op[1], op[2], op[3], op[4], op[5])
    out = eng.get_csr()
    # only the assembly kernels ran (tangent: element matrices + row gather, residual): no pattern or table rebuild
    assert eng.kernel_launches - launches0 == 3
    r = flows.compare(ref, out)
    assert r["pattern_equal"] and r["val_diff"] <= TOL and r["rhs_diff"] <= TOL
    assert not np.array_equal(first[2], out[2])


def test_moving_mesh_update_coords(eng):
    c = flows.build_case("laplace_q1_hex", 5)
    c.run_engine(eng=eng)
    c.coords = np.ascontiguousarray(c.coords * np.array([1.0, 1.1, 0.9]))
    ref = c.run_oracle()
    eng.update_coords(c.coords)
    eng.new_solver(c.n_eqn)
    for op in c.ops:
        if op[0] == "matrix":
            eng.stiffness_matrix_computation(op[1], op[2], op[3], op[4], op[5], incremental=op[6])
        else:
            eng.body_force_computation(op[1], op[2], op[3])
    r = flows.compare(ref, eng.get_csr())
    assert r["pattern_equal"] and r["val_diff"] <= TOL and r["rhs_diff"] <= TOL


def test_generic_kernel_agrees_with_q1_hot_path(eng):
    """quad degree 5 (27 points) takes the generic staged kernel, degree 3 the one-thread-per-element kernel; for a
    trilinear Laplacian on affine elements both integrate exactly."""
    c = flows.build_case("laplace_q1_hex", 6, perturb=False)
    a = c.run_engine(eng=eng)
    c.ops = [("matrix", E.K_LAPLACE, [1.0], 5, 0, 0, True), ("body", [1.0], 3, 0)]
    b = c.run_engine(eng=eng)
    r = flows.compare(a, b)
    assert r["pattern_equal"] and r["val_diff"] <= 1e-11


def test_solver_interface_odd_contributions_and_norm(eng):
    c = flows.build_case("laplace_q1_hex", 4)
    rp, col, val, rhs = c.run_engine(eng=eng)
    assert eng.norm() == pytest.approx(np.linalg.norm(rhs) / len(rhs), rel=1e-13)  # Eigen3.hpp:133-138 quirk
    assert eng.get_value(3) == rhs[3]
    rows = np.array([0, 1]); cols = col[rp[0]:rp[0] + 2]
    # (0, cols) exist; (1, cols) may not: use row 0 only
    eng.insert_to_lhs(np.array([[1.5, -2.0]]), rows[:1], cols)
    eng.insert_to_rhs(np.array([0.25, 0.5]), rows)
    rp2, col2, val2, rhs2 = eng.get_csr()
    assert val2[rp[0]] == val[rp[0]] + 1.5 and val2[rp[0] + 1] == val[rp[0] + 1] - 2.0
    assert rhs2[0] == rhs[0] + 0.25 and rhs2[1] == rhs[1] + 0.5
    with pytest.raises(E.EngineError, match="not been properly set up"):
        far = np.array([c.n_eqn - 1])
        eng.insert_to_lhs(np.array([[1.0]]), rows[:1], far)   # recorded on the device (no wait per call) ...
        eng.finish_assembly()                                  # ... and reported when the assembly is finished
    with pytest.raises(E.EngineError, match="out of bound"):
        eng.insert_to_rhs(np.array([1.0]), np.array([c.n_eqn]))


def test_asynchronous_hand_off_double_buffers_the_system(eng):
    """isl_get_csr_async: values / rhs of step k arrive in pinned host buffers while step k+1 is assembled into the second
    set of device buffers; results equal the blocking isl_get_csr, and the released system cannot be used any more"""
    import torch
    c = flows.build_case("laplace_q1_hex", 9, perturb=False)
    f = c.fields[0]
    eng.set_mesh(c.shape, c.geom_deg, c.coords, c.conn)
    eng.set_field(0, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"], f["eqn"], f["status"], f["presc"], f["values"])
    ref = {}
    for kappa in (1.0, 2.5, 0.5):
        eng.new_solver(c.n_eqn)
        eng.stiffness_matrix_computation(E.K_LAPLACE, [kappa], 3, 0, 0, True)
        eng.body_force_computation([kappa], 3, 0)
        ref[kappa] = eng.get_csr()
    nnz = len(ref[1.0][2])
    bufs = [(torch.empty(nnz, dtype=torch.float64).pin_memory().numpy(), torch.empty(c.n_eqn, dtype=torch.float64).pin_memory().numpy())
            for _ in range(3)]
    for k, kappa in enumerate((1.0, 2.5, 0.5)):
        eng.new_solver(c.n_eqn)
        eng.stiffness_matrix_computation(E.K_LAPLACE, [kappa], 3, 0, 0, True)
        eng.body_force_computation([kappa], 3, 0)
        eng.get_csr_async(*bufs[k])
    with pytest.raises(E.EngineError, match="create a new solver"):
        eng.body_force_computation([1.0], 3, 0)
    eng.copy_wait()
    for k, kappa in enumerate((1.0, 2.5, 0.5)):
        assert np.array_equal(bufs[k][0], ref[kappa][2]) and np.array_equal(bufs[k][1], ref[kappa][3])
    eng.new_solver(c.n_eqn)
    eng.stiffness_matrix_computation(E.K_LAPLACE, [1.0], 3, 0, 0, True)
    eng.body_force_computation([1.0], 3, 0)
    out = eng.get_csr()
    assert np.array_equal(out[2], ref[1.0][2]) and np.array_equal(out[3], ref[1.0][3])


def test_fields_replaced_during_assembly_is_an_error(eng):
    """isl_field_set / isl_mesh_set drop the pattern and the values; when that happens between two assembly
    calls on one solver (a second FieldBinder, a rebuilt binder) the system must not silently continue empty"""
    c = flows.build_case("laplace_q1_hex", 3)
    c.run_engine(eng=eng)
    f = c.fields[0]
    eng.set_field(0, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"], f["eqn"], f["status"], f["presc"], f["values"])
    with pytest.raises(E.EngineError, match="create a new solver"):
        eng.body_force_computation([1.0], 3, 0)
    with pytest.raises(E.EngineError, match="create a new solver"):
        eng.finish_assembly()
    eng.new_solver(c.n_eqn)                      # a fresh solver is fine again
    eng.body_force_computation([1.0], 3, 0)      # right-hand side only: replacing the fields now loses nothing
    eng.set_field(0, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"], f["eqn"], f["status"], f["presc"], f["values"])
    eng.stiffness_matrix_computation(E.K_LAPLACE, [1.0], 3, 0, 0, True)
    eng.finish_assembly()
    with pytest.raises(E.EngineError, match="field index out of range"):
        eng.compute_residual_forces(E.K_LAPLACE, [1.0], 3, 0, 7)


def test_error_behaviour(eng):
    c = flows.build_case("laplace_q1_hex", 3)
    c.run_engine(eng=eng)
    with pytest.raises(E.EngineError, match="HyperElastic kernel"):
        eng.stiffness_matrix_computation(E.K_HYPEL_STVENANT, [1.0, 1.0], 3, 0, 0)
    with pytest.raises(E.EngineError, match="not set"):
        eng.stiffness_matrix_computation(E.K_LAPLACE, [1.0], 3, 1, 1)
    with pytest.raises(E.EngineError, match="unknown kernel"):
        eng.stiffness_matrix_computation(99, [1.0], 3, 0, 0)


def test_large_mesh_properties(eng):
    """64^3 Q1 Laplace (262k elements): size-independent properties instead of the oracle:
    symmetry, zero row sums for interior rows (constants in the kernel), linearity in kappa, and
    total 'energy' 1^T K 1 + lift consistency."""
    from insilico_b200 import meshgen
    n = 64
    coords, conn, _ = meshgen.unit_cube_hex(n, n, n)
    coords = meshgen.perturb_interior(coords, 1.0 / n, 0.15)
    c = flows.Case(E.HEX, 1, coords, conn)
    c.add_field(1, 1, dirichlet=lambda x: 1.0 + 0 * x[:, :1])   # u = 1 on the boundary
    c.ops = [("matrix", E.K_LAPLACE, [1.0], 3, 0, 0, True)]
    rp, col, val, rhs = c.run_engine(eng=eng)
    A = sp.csr_matrix((val, col, rp))
    assert A.shape[0] == (n - 1) ** 3 and A.nnz == (3 * (n - 1) - 2) ** 3
    scale = abs(A).max()
    assert abs(A - A.T).max() <= 1e-12 * scale
    # constant solution u=1: A*1 - rhs = 0 because rhs = -K_ad*1 and full row sums vanish
    assert np.abs(A @ np.ones(A.shape[0]) - rhs).max() <= 1e-11 * scale
    c.ops = [("matrix", E.K_LAPLACE, [3.0], 3, 0, 0, True)]
    rp3, col3, val3, rhs3 = c.run_engine(eng=eng)
    assert np.array_equal(rp, rp3) and np.array_equal(col, col3)
    assert H.csr_rel_diff(rp, 3.0 * val, val3) <= 1e-12 and H.vec_rel_diff(3.0 * rhs, rhs3) <= 1e-12


@pytest.mark.parametrize("env", [{"ISL_Q1_MODE": "atomic"}, {"ISL_Q1_FAST": "0"}, {"ISL_PATCH_ROWS": "64"},
                                 {"ISL_PATCH_ROWS": "200", "ISL_PATCH_THREADS": "256"}, {"ISL_PATCH_ROWS": "360"},
                                 {"ISL_PATCH_ROWS": "512", "ISL_PATCH_THREADS": "256", "ISL_PATCH_CTAS": "1"},
                                 {"ISL_PATCH_WS": "1"}, {"ISL_PATCH_WS": "1", "ISL_PATCH_ROWS": "120"}, {"ISL_Q1_FAST": "3"},
                                 {"ISL_AFFINE_KERNEL": "0"}, {"ISL_AFF_SPLIT": "1", "ISL_AFF_THREADS": "288"},
                                 {"ISL_AFF_THREADS": "320"}, {"ISL_PATCH_CTAS": "3", "ISL_PATCH_ROWS": "260"}])
@pytest.mark.parametrize("name,n,permute", [("laplace_q1_hex", 13, False), ("laplace_q1_hex_values", 9, True)])
def test_q1_hot_path_variants(monkeypatch, env, name, n, permute):
    """one-thread-per-element atomic kernel and shared-memory patch kernel (several patch sizes, Morton ordering of a
    permuted element list) all reproduce the oracle; accumulation into a non-empty matrix (second assembly call in the
    same solver) takes the read-modify-write path of complete rows."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    e = E.Engine(0)
    try:
        c = flows.build_case(name, n, True, permute)
        ref = c.run_oracle()
        out = c.run_engine(eng=e)
        r = flows.compare(ref, out)
        assert r["pattern_equal"] and r["val_diff"] <= TOL and r["rhs_diff"] <= TOL, r
        # assemble the matrix twice into the same solver: K doubles, lift doubles
        ops = c.ops
        c.ops = [ops[0], ops[0]]
        ref2 = c.run_oracle()
        out2 = c.run_engine(eng=e)
        r2 = flows.compare(ref2, out2)
        assert r2["pattern_equal"] and r2["val_diff"] <= TOL and r2["rhs_diff"] <= TOL, r2
    finally:
        e.close()


def test_cpp_facade_example_matches_python_flow(eng):
    """examples/heat_dirichlet.cpp drives the engine through include/insilico_b200.hpp with the reference's call
    sequence; its system equals the one of the Python flow (and therefore the oracle's)."""
    import os
    import subprocess
    exe = os.path.join(H.ROOT, "examples", "heat_dirichlet")
    out = subprocess.run([exe, "7"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    ndof, nnz, norm, total = out.stdout.split()
    c = flows.build_case("laplace_q1_hex", 7, perturb=False)
    c.ops = c.ops[:1]
    rp, col, val, rhs = c.run_engine(eng=eng)
    ref = c.run_oracle()
    assert int(ndof) == len(rhs) and int(nnz) == len(val)
    assert float(norm) == pytest.approx(np.linalg.norm(ref[3]) / len(rhs), rel=1e-12)
    assert float(total) == pytest.approx(ref[2].sum(), rel=1e-9, abs=1e-9 * np.abs(ref[2]).max())


def test_q1_non_lattice_connectivity_falls_back_to_atomic_kernel(eng):
    """Rotating the local node order of some hexahedra (still valid elements) breaks the lattice property the patch
    kernel relies on (a node must be local node a of at most one element); the engine detects that in the patch
    preprocessing and uses the one-thread-per-element kernel.  Result must still equal the oracle."""
    c = flows.build_case("laplace_q1_hex", 9, perturb=True)
    rot = np.array([1, 2, 3, 0, 5, 6, 7, 4])          # rotation about the local zeta axis
    conn = c.conn.copy()
    conn[::3] = conn[::3][:, rot]
    c2 = flows.Case(E.HEX, 1, c.coords, conn)
    c2.add_field(1, 1, dirichlet=lambda x: H.fund_sol_laplace(x, np.full(3, -0.5)))
    c2.ops = c.ops
    ref = c2.run_oracle()
    out = c2.run_engine(eng=eng)
    r = flows.compare(ref, out)
    assert r["pattern_equal"] and r["val_diff"] <= TOL and r["rhs_diff"] <= TOL, r
    # same operator as the unrotated mesh (local numbering does not change the assembled system)
    r0 = flows.compare(c.run_oracle(), out)
    assert r0["pattern_equal"] and r0["val_diff"] <= 1e-11


@pytest.mark.parametrize("n", [1, 2, 3])
def test_degenerate_sizes(eng, n):
    """n=1: every node is on the boundary (no equation at all, nnz = 0); n=2: one ACTIVE node (1x1 system whose rhs is
    pure Dirichlet lift); n=3: 8 equations.  Same flow as the big cases."""
    c = flows.build_case("laplace_q1_hex", n, perturb=(n > 1))
    ref = c.run_oracle()
    out = c.run_engine(eng=eng)
    assert len(out[3]) == (n - 1) ** 3
    assert np.array_equal(ref[0], out[0]) and np.array_equal(ref[1], out[1])
    if n > 1:
        r = flows.compare(ref, out)
        assert r["val_diff"] <= TOL and r["rhs_diff"] <= TOL


def test_inactive_dofs_are_skipped(eng):
    """INACTIVE DoFs (immersed methods, base/dof/DegreeOfFreedom.hpp:33-38) neither get an equation nor a lift; elements
    whose DoFs are all INACTIVE contribute nothing (collectFromDoFs.hpp:118-134)."""
    c = flows.build_case("laplace_q1_hex", 5, perturb=True)
    f = c.fields[0]
    inactive = (f["status"][:, 0] == 0) & (c.coords[:, 0] < 0.45)
    f["status"][inactive, 0] = E.INACTIVE
    f["eqn"], c.n_eqn = E.number_dofs_consecutively(f["status"])
    ref = c.run_oracle()
    out = c.run_engine(eng=eng)
    r = flows.compare(ref, out)
    assert r["pattern_equal"] and r["val_diff"] <= TOL and r["rhs_diff"] <= TOL, r
    assert c.n_eqn < 4 ** 3


def test_affine_kernel_multi_patch_equals_general_kernel(monkeypatch):
    """unperturbed 40^3 mesh (all elements affine, ~160 patches): the all-affine kernel and the general patch kernel
    (ISL_AFFINE_KERNEL=0) produce the same system; a permuted 14^3 mesh is also compared with the oracle."""
    c = flows.build_case("laplace_q1_hex", 14, False, True)
    ref = c.run_oracle()
    big = flows.build_case("laplace_q1_hex", 40, False, False)
    outs = []
    for knob in ("1", "0"):
        monkeypatch.setenv("ISL_AFFINE_KERNEL", knob)
        e = E.Engine(0)
        try:
            r = flows.compare(ref, c.run_engine(eng=e))
            assert r["pattern_equal"] and r["val_diff"] <= TOL and r["rhs_diff"] <= TOL, r
            outs.append(big.run_engine(eng=e))
        finally:
            e.close()
    r = flows.compare(outs[0], outs[1])
    assert r["pattern_equal"] and r["val_diff"] <= TOL and r["rhs_diff"] <= TOL, r


@pytest.mark.parametrize("threads,rows", [("128", "100"), ("256", "100"), ("128", "256"), ("256", "37")])
def test_rowgather_kernel_equals_oracle(monkeypatch, threads, rows):
    """isl_rowgather.cuh on the device (the same routines pass tests/test_rowgather_emu.py on the host): stencil-sum
    kernel on affine meshes, gathered local matrices on perturbed ones, store and accumulate mode of the bulk write-out."""
    monkeypatch.setenv("ISL_Q1_ROWS", "1")
    monkeypatch.setenv("ISL_ROWS_THREADS", threads)
    monkeypatch.setenv("ISL_PATCH_ROWS", rows)
    e = E.Engine(0)
    try:
        for name, n, perturb, permute in [("laplace_q1_hex", 13, False, False), ("laplace_q1_hex_values", 9, False, True),
                                          ("laplace_q1_hex", 11, True, True)]:   # perturbed: k_q1hex_rows_general
            c = flows.build_case(name, n, perturb, permute)
            ref = c.run_oracle()
            r = flows.compare(ref, c.run_engine(eng=e))
            assert r["pattern_equal"] and r["val_diff"] <= TOL and r["rhs_diff"] <= TOL, r
            ops = c.ops
            c.ops = [ops[0], ops[0]]     # second assembly accumulates into complete rows
            r2 = flows.compare(c.run_oracle(), c.run_engine(eng=e))
            assert r2["pattern_equal"] and r2["val_diff"] <= TOL and r2["rhs_diff"] <= TOL, r2
    finally:
        e.close()


def test_neumann_with_explicit_rows_equals_field_form(eng):
    """isl_assemble_neumann_rows (equation numbers per surface element, what the reference-tree binding passes) adds the
    same right-hand side as the field form"""
    c = flows.build_case("neumann_q1_hex", 5, True, False)
    a = c.run_engine(eng=eng)
    f = c.fields[0]
    op = [o for o in c.ops if o[0] == "neumann"][0]
    de, sx, sp = c.surface_elements(op, E)
    mode, data = c.surface_force(op, sx, E.surface_points)
    rows = np.where(f["status"] == 0, f["eqn"], -1)[f["elem_dof"][de]].reshape(len(de), -1)
    c.ops = [o for o in c.ops if o[0] != "neumann"]
    eng.set_mesh(c.shape, c.geom_deg, c.coords, c.conn)
    eng.set_field(0, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"], f["eqn"], f["status"], f["presc"], f["values"])
    eng.new_solver(c.n_eqn)
    for o in c.ops:
        if o[0] == "matrix":
            eng.stiffness_matrix_computation(o[1], o[2], o[3], o[4], o[5], incremental=o[6])
        else:
            eng.body_force_computation(o[1], o[2], o[3])
    eng.neumann_force_computation_rows(c.shape, c.geom_deg, sx, sp, op[2], f["fe_deg"], f["ds"], rows, mode, data)
    b = eng.get_csr()
    r = flows.compare(a, b)
    assert r["pattern_equal"] and r["val_diff"] <= TOL and r["rhs_diff"] <= TOL, r


@pytest.mark.parametrize("shape,fe_deg,n,permute", [(E.HEX, 2, 4, True), (E.HEX, 3, 3, False), (E.TET, 2, 4, True), (E.QUAD, 2, 7, True),
                                                     (E.QUAD, 3, 5, False), (E.TRI, 2, 6, True), (E.HEX, 1, 5, True), (E.TET, 1, 3, False)])
def test_device_dof_generation_equals_host(eng, shape, fe_deg, n, permute):
    """isl_dof_generate_device (radix sorts + scans) gives the ids of the host function, i.e. of base::dof::generate
    (pinned against the reference run in test_reference_run.py), also on a permuted element order"""
    coords, conn = flows.make_mesh(shape, n, True, permute)
    eng.set_mesh(shape, 1, coords, conn)
    ed_dev, n_dev = eng.dof_generate(fe_deg)
    ed, nobj = E.dof_generate(shape, 1, conn, fe_deg)
    assert n_dev == nobj
    assert np.array_equal(ed_dev, ed)
