"""The benchmark workload at BASELINE's full size (256^3 Q1-hex Laplace, stiffness + RHS): every matrix and rhs entry
against the oracle through the exact scaling argument of tests/structured_check.py."""
import numpy as np
import pytest

from tests import flows, structured_check as SC


def bench_dirichlet(x):
    d = np.sqrt(((x + 0.5) ** 2).sum(axis=1))
    return (1.0 / (4.0 * np.pi)) / d


def test_scaling_argument_holds_for_the_oracle():
    """CPU: the oracle's own 16^3 system equals the 8^3 stencil scaled by 1/2 up to the summation order of the element
    contributions (a few ulp)"""
    S0, body0, n0 = SC.oracle_stencil(8)
    from insilico_b200 import engine as E
    c = flows.build_case("laplace_q1_hex", 16, perturb=False)
    c.ops = [("matrix", E.K_LAPLACE, [1.0], 3, 0, 0, True), ("body", [1.0], 3, 0)]
    rp, col, val, rhs = c.run_oracle()
    from tests import helpers as H
    val_err, rhs_err = SC.check_structured_system(16, rp, col, val, rhs, lambda x: H.fund_sol_laplace(x, np.full(3, -0.5)),
                                                  S0, body0, n0)
    assert val_err <= 5e-15 and rhs_err <= 5e-15


def test_bench_workload_matches_the_checker_numbering():
    """CPU: bench.py's workload builder (partition.structured_laplace_slab) assembled by the oracle passes the checker,
    i.e. its equation numbering is the x-fastest interior numbering the checker decodes"""
    from insilico_b200 import partition
    from oracle import oracle as orc
    n = 12
    wl = partition.structured_laplace_slab(n, n, n, 0, 1, bench_dirichlet)
    prob = orc.Problem(orc.HEX, 1, wl["coords"], wl["conn"].astype(np.int64))
    prob.set_field(0, 1, 1, wl["n_obj"], wl["elem_dof"].astype(np.int64), wl["eqn"], wl["status"], wl["presc"], wl["values"])
    s = orc.System(wl["n_eqn_local"])
    s.stiffness(prob, orc.K_LAPLACE, [1.0], 3, 0, 0, incremental=True, nthreads=1)
    s.bodyforce(prob, [1.0], 3, 0)
    rp, col, val, rhs = s.finish()
    S0, body0, n0 = SC.oracle_stencil(8)
    # 12 is not a power of two times 8: the scaling is then exact only up to rounding
    SC.check_structured_system(n, rp, col, val, rhs, bench_dirichlet, S0, body0, n0, tol=1e-13)


def test_checker_detects_a_wrong_entry():
    S0, body0, n0 = SC.oracle_stencil(8)
    from insilico_b200 import engine as E
    from tests import helpers as H
    c = flows.build_case("laplace_q1_hex", 8, perturb=False)
    c.ops = [("matrix", E.K_LAPLACE, [1.0], 3, 0, 0, True), ("body", [1.0], 3, 0)]
    rp, col, val, rhs = c.run_oracle()
    val = val.copy(); val[len(val) // 3] *= 1.0 + 1e-9
    with pytest.raises(AssertionError):
        SC.check_structured_system(8, rp, col, val, rhs, lambda x: H.fund_sol_laplace(x, np.full(3, -0.5)), S0, body0, n0)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [32, 256])
def test_engine_full_size_structured_system_equals_scaled_oracle_stencil(n):
    import psutil
    if n == 256 and psutil.virtual_memory().available < 40e9:
        pytest.skip("needs about 30 GB of host memory for the 4.5e8-entry system and the checks")
    from insilico_b200 import engine as E
    from insilico_b200 import partition
    S0, body0, n0 = SC.oracle_stencil(8)
    wl = partition.structured_laplace_slab(n, n, n, 0, 1, bench_dirichlet)       # exactly bench.py's workload
    eng = E.Engine(0)
    try:
        eng.set_mesh(E.HEX, 1, wl["coords"], wl["conn"])
        eng.set_field(0, 1, 1, wl["n_obj"], wl["elem_dof"], wl["eqn"], wl["status"], wl["presc"], wl["values"])
        eng.new_solver(wl["n_eqn_local"])
        eng.register_fields(0, 0)
        for _ in range(2):                                                        # second pass: pattern and patches cached
            eng.new_solver(wl["n_eqn_local"])
            eng.stiffness_matrix_computation(E.K_LAPLACE, [1.0], 3, 0, 0, True)
            eng.body_force_computation([1.0], 3, 0)
        rp, col, val, rhs = eng.get_csr()
    finally:
        eng.close()
    del wl
    SC.check_structured_system(n, rp, col, val, rhs, bench_dirichlet, S0, body0, n0)


def _perturbed_workload(n):
    from insilico_b200 import engine as E
    from insilico_b200 import meshgen
    coords, conn, _ = meshgen.unit_cube_hex(n, n, n)
    coords = meshgen.perturb_interior(coords, 1.0 / n, 0.15)
    c = flows.Case(E.HEX, 1, coords, conn)
    c.add_field(1, 1, dirichlet=lambda x: 1.0 + 0 * x[:, :1])
    c.ops = [("matrix", E.K_LAPLACE, [1.0], 3, 0, 0, True)]
    return c


def test_sub_box_argument_holds_for_the_oracle():
    """CPU: the perturbed-mesh checker (pattern, A 1 = rhs, symmetry, oracle on sub-boxes) accepts the oracle's system"""
    n = 12
    c = _perturbed_workload(n)
    rp, col, val, rhs = c.run_oracle()
    err = SC.check_perturbed_system(n, c.coords, rp, col, val, rhs, boxes=4, box=4)
    assert err <= 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("n", [24, 256])
def test_engine_full_size_perturbed_mesh_properties_and_sub_box_oracle_parity(n):
    import psutil
    if n == 256 and psutil.virtual_memory().available < 48e9:
        pytest.skip("needs about 40 GB of host memory")
    c = _perturbed_workload(n)
    rp, col, val, rhs = c.run_engine()
    SC.check_perturbed_system(n, c.coords, rp, col, val, rhs)


# ---- BASELINE configs 3 and 4 at their benchmark sizes: oracle parity on sampled sub-meshes ---------------------------
def _oracle_system(w):
    from oracle import oracle as orc
    prob = orc.Problem(w.shape, w.geom_deg, w.coords, w.conn.astype(np.int64))
    for i, f in enumerate(w.fields):
        prob.set_field(i, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"].astype(np.int64), f["eqn"], f["status"], f["presc"], f["values"])
    s = orc.System(w.n_eqn)
    for op in w.ops:
        if op[0] == "matrix":
            s.stiffness(prob, op[1], op[2], op[3], op[4], op[5], incremental=op[6])
    return s.finish()


@pytest.mark.parametrize("cfg,n,hw", [("C3", 4, 0.3), ("C4", 4, 0.3), ("C5", 4, 0.3)])
def test_sub_mesh_argument_holds_for_the_oracle(cfg, n, hw):
    """CPU: the sub-mesh checker accepts the oracle's own full system (and so checks nothing but the argument)"""
    from insilico_b200 import workloads
    from tests import subbox_check as SB
    w = workloads.build(cfg, n)
    rp, col, val, rhs = _oracle_system(w)
    err, compared = SB.check_subboxes(w, rp, col, val, n_boxes=2, half_width=hw)
    assert err <= 1e-14 and compared > 50
    val2 = val.copy()
    val2[len(val2) // 2] += 1e-6 * np.abs(val).max()     # the checker can fail: some box must see a perturbed entry
    with pytest.raises(AssertionError):
        for seed in range(40):
            SB.check_subboxes(w, rp, col, val2, n_boxes=2, half_width=hw, seed=seed)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,n,hw", [("C3", 12, 0.13), ("C4", 12, 0.1), ("C5", 12, 0.1), ("C3", 64, 0.025), ("C4", 64, 0.02)])
def test_engine_config_sizes_sub_mesh_oracle_parity(cfg, n, hw):
    """bench.py --config C3 / C4 (/ C5) workloads on the engine, at the benchmark size for C3 and C4: entries between DoFs
    that only the elements of a sampled box touch equal the oracle's assembly of that box"""
    import psutil
    from insilico_b200 import engine as E
    from insilico_b200 import workloads
    from tests import subbox_check as SB
    if n >= 64 and psutil.virtual_memory().available < 64e9:
        pytest.skip("needs about 50 GB of host memory")
    w = workloads.build(cfg, n)
    eng = E.Engine(0)
    try:
        w.upload(eng)
        eng.new_solver(w.n_eqn)
        w.register(eng)
        w.step(eng)
        rp, col, val, rhs = eng.get_csr()
    finally:
        eng.close()
    err, compared = SB.check_subboxes(w, rp, col, val, n_boxes=3, half_width=hw)
    assert compared > 100
