"""General linear constraints: slave DoF components u = rhs + sum_j w_j u_master_j (base/dof/Constraint.hpp:57-140),
collected by asmb::collectFromDoFs (base/asmb/collectFromDoFs.hpp:112-131) and applied by asmb::assembleMatrix /
assembleForces (base/asmb/assembleMatrix.hpp:56-130,212-338, assembleForces.hpp:58-139): weighted extra rows and
columns for the masters, prescribed part lifted to the rhs.

Fixtures tests/golden/refrun/*_linear_*.npz come from the unmodified reference (tools/make_ref_goldens.py).  The file
is named test_zz_* so that it runs last: its GPU tests exercise isl_field_set_constraints, which was written in a
session without GPU minutes (verified there against the oracle only through the mock ABI)."""
import os

import numpy as np
import pytest

from tests import flows
from tests.test_reference_run import (APPS_B200, GOLD_LINEAR, _check_numbering, _driver_on_binding, _load)

IDS = [os.path.basename(p)[:-4] for p in GOLD_LINEAR]


def test_fixtures_hold_master_slave_constraints():
    assert len(GOLD_LINEAR) >= 6
    for p in GOLD_LINEAR:
        _, case = _load(p)
        assert any(f["linear"] for f in case.fields)
        for f in case.fields:
            for obj, comp, rhs, masters in f["linear"]:
                assert f["status"][obj, comp] == 1 and f["eqn"][obj, comp] < 0 and len(masters) >= 2


@pytest.mark.parametrize("path", GOLD_LINEAR, ids=IDS)
def test_oracle_reproduces_reference_run_with_linear_constraints(path):
    g, case = _load(path)
    _check_numbering(g, case)
    res = flows.compare((g["rowptr"], g["col"], g["val"], g["rhs"]), case.run_oracle(register=bool(g["register"])))
    assert res["pattern_equal"] and res["val_diff"] <= 1e-13 and res["rhs_diff"] <= 1e-13, res


@pytest.mark.skipif(not os.path.isdir(APPS_B200), reason="oracle/_ref/apps_b200 not built (needs /root/reference)")
@pytest.mark.parametrize("path", GOLD_LINEAR, ids=IDS)
def test_reference_api_on_binding_with_mock_abi_linear_constraints(tmp_path, path):
    res = _driver_on_binding(path, "ref_driver_mock", tmp_path)
    assert res["val_diff"] <= 1e-13 and res["rhs_diff"] <= 1e-13, res


def test_constraint_errors_are_reported():
    """argument checks of isl_field_set_constraints that need no GPU are made by the host wrapper's caller contract:
    masters must be ACTIVE (flows.Case.constraint_arrays asserts it)"""
    c = flows.build_case("laplace_q1_hex_linear", 4, True, False)
    f = c.fields[0]
    con_dof, con_ptr, meq, w = c.constraint_arrays(f)
    assert len(con_dof) == len(f["linear"]) and con_ptr[-1] == len(meq) == len(w) and (meq >= 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD_LINEAR, ids=IDS)
def test_engine_reproduces_reference_run_with_linear_constraints(path):
    g, case = _load(path)
    _check_numbering(g, case)
    res = flows.compare((g["rowptr"], g["col"], g["val"], g["rhs"]), case.run_engine(register=bool(g["register"])))
    assert res["pattern_equal"], "CSR pattern differs from the reference's finished matrix"
    assert res["val_diff"] <= 1e-12 and res["rhs_diff"] <= 1e-12, res


@pytest.mark.gpu
@pytest.mark.parametrize("name,n", [("laplace_q1_hex_linear", 7), ("laplace_q2_hex_linear", 3), ("laplace_p1_tet_linear", 6),
                                    ("stvenant_q1_hex_linear", 5), ("stokes_p2p1_tet_linear", 3)])
def test_engine_equals_oracle_with_linear_constraints(name, n):
    for register in (False, True):
        res = flows.run_case(name, n=n, perturb=True, register=register)
        assert res["pattern_equal"] and res["val_diff"] <= 1e-12 and res["rhs_diff"] <= 1e-12, (register, res)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(APPS_B200), reason="oracle/_ref/apps_b200 not built (needs /root/reference)")
@pytest.mark.parametrize("path", GOLD_LINEAR, ids=IDS)
def test_reference_api_on_b200_engine_linear_constraints(tmp_path, path):
    _driver_on_binding(path, "ref_driver", tmp_path)


@pytest.mark.gpu
def test_constraint_argument_checks_on_engine():
    from insilico_b200 import engine as E
    c = flows.build_case("laplace_q1_hex_linear", 4, True, False)
    f = c.fields[0]
    eng = E.Engine(0)
    try:
        eng.set_mesh(c.shape, c.geom_deg, c.coords, c.conn)
        eng.set_field(0, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"], f["eqn"], f["status"], f["presc"], f["values"])
        con_dof, con_ptr, meq, w = c.constraint_arrays(f)
        active = int(np.argwhere(f["status"][:, 0] == 0)[0, 0])
        with pytest.raises(RuntimeError, match="CONSTRAINED"):
            eng.set_field_constraints(0, [active], [0, 1], [0], [1.0])
        with pytest.raises(RuntimeError, match="ACTIVE"):
            eng.set_field_constraints(0, con_dof[:1], [0, 1], [-1], [1.0])
        eng.set_field_constraints(0, con_dof, con_ptr, meq, w)
    finally:
        eng.close()
