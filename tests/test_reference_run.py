"""Parity against outputs of the reference ITSELF.

tests/golden/refrun/*.npz were produced by tools/make_ref_goldens.py from oracle/_ref/ref_driver = the unmodified
headers of /root/reference (base::asmb::stiffnessMatrixComputation / computeResidualForces / bodyForceComputation into
base::solver::Eigen3, finishAssembly, debugLHS/debugRHS) compiled against the std-only Boost/Eigen stand-ins of
oracle/compat.  They hold the reference's DoF numbering (element -> DoF ids, status, equation numbers) and its
finished system for 29 cases (6 with general linear constraints and 3 with body forces f(x): see test_zz_linear_constraints.py).

  * CPU (`not gpu`): the oracle restatement reproduces them (numbering/pattern exact, values 1e-13) -> the oracle is
    pinned entry-wise, not only by the reference's 6-digit goldens.
  * where /root/reference exists: the stand-ins themselves are pinned by re-running the reference's own regression
    applications and comparing with the reference's golden files (byte-identical where the reference prints
    deterministic digits).
  * GPU (`gpu`): the CUDA engine, through the C ABI, against the same fixtures (1e-12).
"""
import glob
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import flows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALL_GOLD = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "refrun", "*.npz")))
# the cases with general linear constraints and with general body forces f(x) live in tests/test_zz_linear_constraints.py
# (engine entry points written without GPU minutes: they run last)
_late = lambda p: "_linear" in os.path.basename(p) or "_bodyfun" in os.path.basename(p)
GOLD = [p for p in ALL_GOLD if not _late(p)]
GOLD_LINEAR = [p for p in ALL_GOLD if _late(p)]
REFERENCE = "/root/reference"
APPS = os.path.join(ROOT, "oracle", "_ref", "apps")


def _load(path):
    g = np.load(path, allow_pickle=False)
    case = flows.build_case(str(g["case"]), int(g["n"]), bool(g["perturb"]), bool(g["permute"]))
    return g, case


def _check_numbering(g, case):
    for i, f in enumerate(case.fields):
        assert np.array_equal(g["elem_dof%d" % i], f["elem_dof"]), "element -> DoF ids differ from the reference"
        assert np.array_equal(g["status%d" % i], f["status"]), "DoF status differs from the reference"
        act = f["status"] == 0
        assert np.array_equal(g["eqn%d" % i][act], f["eqn"][act]), "equation numbers differ from the reference"


def test_fixtures_present_and_nontrivial():
    assert len(GOLD) >= 20
    for p in GOLD:
        g = np.load(p)
        assert len(g["val"]) == g["rowptr"][-1] > 0
        assert np.abs(g["val"]).max() > 0 and np.abs(g["rhs"]).max() > 0, p


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_reference_run(path):
    g, case = _load(path)
    _check_numbering(g, case)
    out = case.run_oracle(register=bool(g["register"]))
    res = flows.compare((g["rowptr"], g["col"], g["val"], g["rhs"]), out)
    assert res["pattern_equal"], "CSR pattern differs from the reference's finished matrix"
    assert res["val_diff"] <= 1e-13 and res["rhs_diff"] <= 1e-13, res


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_engine_reproduces_reference_run(path):
    g, case = _load(path)
    _check_numbering(g, case)
    out = case.run_engine(register=bool(g["register"]))
    res = flows.compare((g["rowptr"], g["col"], g["val"], g["rhs"]), out)
    assert res["pattern_equal"], "CSR pattern differs from the reference's finished matrix"
    assert res["val_diff"] <= 1e-12 and res["rhs_diff"] <= 1e-12, res


# ---- the stand-ins are pinned by the reference's own regression applications (only where the reference exists) ------
needs_ref = pytest.mark.skipif(not (os.path.isdir(REFERENCE) and os.path.isdir(APPS)),
                               reason="/root/reference or oracle/_ref/apps not present (GPU box)")


def _run(tmp_path, exe, args, inputs):
    for src in inputs:
        os.symlink(src, os.path.join(tmp_path, os.path.basename(src)))
    return subprocess.run([os.path.join(APPS, exe)] + args, cwd=tmp_path, check=True, capture_output=True, text=True,
                          timeout=600).stdout


@needs_ref
@pytest.mark.parametrize("deg", [1, 2, 3])
def test_reference_app_dofhandler_sparsity_golden(tmp_path, deg):
    d = os.path.join(REFERENCE, "reference", "03-doFHandler")
    _run(str(tmp_path), "doFHandler%d" % deg, ["square_20.smf"], [os.path.join(d, "square_20.smf")])
    mine = open(os.path.join(str(tmp_path), "sparsity.%d.dat" % deg)).read().splitlines()
    gold = open(os.path.join(d, "sparsity.%d.ref.dat" % deg)).read().splitlines()
    assert mine[1:] == gold[1:]  # line 0 names the executable


@needs_ref
def test_reference_app_areavolume_golden(tmp_path):
    d = os.path.join(REFERENCE, "reference", "02-areaVolume")
    out = _run(str(tmp_path), "areaVolume", [], [os.path.join(d, f) for f in
                                                 ("input.dat", "sphere.coords", "sphere.tetrahedron.conn",
                                                  "sphere.tetrahedron.smf")])
    assert out == open(os.path.join(d, "measure.ref.dat")).read()


@needs_ref
@pytest.mark.parametrize("dim,meshes", [(2, ["quad.002", "quad.005", "quad.010", "quad.020", "quad.040"]),
                                        (3, ["cube.002", "cube.004", "cube.008", "cube.012"])])
def test_reference_app_linear_elastic_golden(tmp_path, dim, meshes):
    d = os.path.join(REFERENCE, "reference", "06-elastic")
    gold = dict(l.split() for l in open(os.path.join(d, "linearElastic%dD.ref.dat" % dim)) if not l.startswith("#"))
    for m in meshes:
        out = _run(str(tmp_path), "linearElastic%dD" % dim, [m + ".smf"], [os.path.join(d, m + ".smf")])
        assert out.split()[0] == gold[m[5:]], (m, out)


@needs_ref
def test_reference_app_compressible_newton_history_golden(tmp_path):
    """HyperElastic<NeoHookeanCompressible> tangent + residual, displacement controlled: every printed |F| and |x| of
    the Newton history equals the golden to its 6 digits, except residual norms below 1e-11 (converged iterates, i.e.
    solver rounding noise, where the reference's own CG and the stand-in's differ)."""
    d = os.path.join(REFERENCE, "reference", "06-elastic")
    out = _run(str(tmp_path), "compressible", ["quad.020.smf", "inputCompRefD.dat"],
               [os.path.join(d, "quad.020.smf"), os.path.join(d, "inputCompRefD.dat")])
    mine = [l.split() for l in out.splitlines() if l.strip() and not l.startswith("#")]
    gold = [l.split() for l in open(os.path.join(d, "compRefOutD.dat")) if l.strip() and not l.startswith("#")]
    assert len(mine) == len(gold) > 20
    compared = 0
    for a, b in zip(mine, gold):
        assert a[:2] == b[:2] and len(a) == len(b)
        for x, y in zip(a[2:], b[2:]):
            if float(y) < 1e-11:
                assert float(x) < 1e-11
            else:
                assert x == y, (a, b)
                compared += 1
    assert compared >= 30


# ---- the reference's applications, unmodified, on the B200 binding --------------------------------------------------
# oracle/_ref/apps_b200/<app>       = application + include/insilico_b200_reference.hpp + libinsilico_b200.so (GPU)
# oracle/_ref/apps_b200/<app>_mock  = same object file + the oracle-backed mock of the C ABI (CPU check of the binding)
# expected output = what the application prints with the reference's own base::solver::Eigen3
# (tests/golden/refrun_apps, tools/make_ref_app_goldens.py)
from tests import ref_apps_cases as RA  # noqa: E402

APPS_B200 = os.path.join(ROOT, "oracle", "_ref", "apps_b200")


def _run_binding_app(name, suffix, tmp_path):
    exe, args = RA.prepare(name, str(tmp_path))
    path = os.path.join(APPS_B200, exe + suffix)
    p = subprocess.run([path] + args, cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    out = p.stdout
    if RA.output_file(name):
        out += open(os.path.join(str(tmp_path), RA.output_file(name))).read()
    expected = open(os.path.join(ROOT, "tests", "golden", "refrun_apps", name + ".out")).read()
    assert RA.same_output(out, expected), "\n" + out[-3000:] + "\n--- expected ---\n" + expected[-3000:]


@pytest.mark.skipif(not os.path.isdir(APPS_B200), reason="oracle/_ref/apps_b200 not built (needs /root/reference)")
@pytest.mark.parametrize("name", sorted(RA.CASES))
def test_unmodified_reference_app_on_binding_with_mock_abi(tmp_path, name):
    _run_binding_app(name, "_mock", tmp_path)


@pytest.mark.skipif(not os.path.isdir(APPS_B200), reason="oracle/_ref/apps_b200 not built (needs /root/reference)")
def test_binding_app_without_gpu_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe, args = RA.prepare("dirichlet_tet6", str(tmp_path))
    p = subprocess.run([os.path.join(APPS_B200, exe)] + args, cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(APPS_B200), reason="oracle/_ref/apps_b200 not built (needs /root/reference)")
@pytest.mark.parametrize("name", sorted(n for n in RA.CASES if n not in RA.LATE))
def test_unmodified_reference_app_on_b200_engine(tmp_path, name):
    _run_binding_app(name, "", tmp_path)


# ---- every fixture case through the binding: reference API -> flattening -> C ABI -----------------------------------
def _driver_on_binding(path, exe, tmp_path):
    from tools import make_ref_goldens as G
    g = np.load(path, allow_pickle=False)
    spec = [c for c in G.CASES if c[0] == os.path.basename(path)[:-4]][0]
    case = flows.build_case(spec[1], spec[3], spec[4], spec[5])
    old = G.DRIVER
    G.DRIVER = os.path.join(APPS_B200, exe)
    try:
        r = G.run_reference(case, spec[2], spec[6], str(tmp_path))
    finally:
        G.DRIVER = old
    for i in range(len(case.fields)):
        assert np.array_equal(r["elem_dof%d" % i], g["elem_dof%d" % i])
        assert np.array_equal(r["status%d" % i], g["status%d" % i])
    res = flows.compare((g["rowptr"], g["col"], g["val"], g["rhs"]), (r["rowptr"], r["col"], r["val"], r["rhs"]))
    assert res["pattern_equal"], "CSR pattern differs from the reference's finished matrix"
    # material constants are recovered by probing (1e-15), hence 1e-13 instead of bit-level agreement
    assert res["val_diff"] <= 1e-12 and res["rhs_diff"] <= 1e-12, res
    return res


def _two_devices_one_thread(exe, tmp_path, extra_env):
    """SURVEY 8(b) threading row: ONE host thread drives two devices through the binding (B200::selectDevice), the
    assembly calls of two solvers interleaved call by call; both finished systems equal the reference run's"""
    from tools import make_ref_goldens as G
    path = [p for p in GOLD if os.path.basename(p).startswith("stokes_p2p1_tet")][0]
    g = np.load(path, allow_pickle=False)
    spec = [c for c in G.CASES if c[0] == os.path.basename(path)[:-4]][0]
    case = flows.build_case(spec[1], spec[3], spec[4], spec[5])
    old, old_env = G.DRIVER, dict(os.environ)
    G.DRIVER = os.path.join(APPS_B200, exe)
    os.environ.update(dict(ISL_DRIVER_DEVICES="2", **extra_env))
    try:
        r0 = G.run_reference(case, spec[2], spec[6], str(tmp_path))
    finally:
        G.DRIVER = old
        os.environ.clear()
        os.environ.update(old_env)
    assert "two devices, one host thread: solvers on devices 0 and 1" in r0["log"]
    out = os.path.join(str(tmp_path), "out")
    for ext in (".lhs.txt", ".rhs.txt"):       # second device: same reader on the second dump
        os.replace(out + ".dev1" + ext, out + ext)
    r1 = G.read_system(out)
    for r in (r0, r1):
        res = flows.compare((g["rowptr"], g["col"], g["val"], g["rhs"]), (r["rowptr"], r["col"], r["val"], r["rhs"]))
        assert res["pattern_equal"] and res["val_diff"] <= 1e-12 and res["rhs_diff"] <= 1e-12, res


@pytest.mark.skipif(not os.path.isdir(APPS_B200), reason="oracle/_ref/apps_b200 not built (needs /root/reference)")
def test_one_host_thread_drives_two_devices_with_mock_abi(tmp_path):
    _two_devices_one_thread("ref_driver_mock", tmp_path, {})


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(APPS_B200), reason="oracle/_ref/apps_b200 not built (needs /root/reference)")
def test_one_host_thread_drives_two_devices_on_b200(tmp_path):
    """two engines with their own streams, systems and binder copies, driven call by call from one thread; both logical
    devices are folded onto GPU 0 (ISL_B200_PHYSICAL_DEVICES=1), which is what a one-GPU test box can run.  Set
    ISL_TEST_PHYSICAL_DEVICES=2 on a box with two GPUs to put the second engine on GPU 1."""
    _two_devices_one_thread("ref_driver", tmp_path, {"ISL_B200_PHYSICAL_DEVICES": os.environ.get("ISL_TEST_PHYSICAL_DEVICES", "1")})


@pytest.mark.skipif(not os.path.isdir(APPS_B200), reason="oracle/_ref/apps_b200 not built (needs /root/reference)")
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_reference_api_on_binding_with_mock_abi(tmp_path, path):
    res = _driver_on_binding(path, "ref_driver_mock", tmp_path)
    assert res["val_diff"] <= 1e-13 and res["rhs_diff"] <= 1e-13, res


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(APPS_B200), reason="oracle/_ref/apps_b200 not built (needs /root/reference)")
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_reference_api_on_b200_engine(tmp_path, path):
    _driver_on_binding(path, "ref_driver", tmp_path)


# ---- no silent host path: the applications must reach the engine's assembly entry points -----------------------------
@pytest.mark.skipif(not os.path.isdir(APPS_B200), reason="oracle/_ref/apps_b200 not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["dirichlet_tet6", "linearElastic3D_cube004", "compressible_quad010", "compressibleWithDriver_quad010"])
def test_applications_use_the_assembly_entry_points_not_insert(tmp_path, name):
    """ISL_MOCK_TRACE counts the ABI calls: the element matrices must come from isl_assemble_matrix; isl_insert_lhs
    (the reference's own host element loop feeding the solver) must not be used at all."""
    import re
    exe, args = RA.prepare(name, str(tmp_path))
    p = subprocess.run([os.path.join(APPS_B200, exe + "_mock")] + args, cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, ISL_MOCK_TRACE="1"))
    assert p.returncode == 0, p.stderr[-2000:]
    m = re.search(r"assemble_matrix (\d+)\s+assemble_residual (\d+)\s+assemble_bodyforce (\d+)\s+insert_lhs (\d+)\s+insert_rhs (\d+)", p.stderr)
    assert m, p.stderr[-500:]
    n_matrix, n_res, n_body, n_ins_lhs, n_ins_rhs = map(int, m.groups())
    assert n_matrix >= 1 and n_ins_lhs == 0 and n_ins_rhs == 0, m.group(0)


@pytest.mark.skipif(not os.path.isdir(APPS_B200), reason="oracle/_ref/apps_b200 not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["mixedPoisson_square020", "mixedPoissonWithDriver_square020", "compressible_neumann_quad010"])
def test_applications_with_surface_forces_use_the_neumann_entry_point(tmp_path, name):
    """base::asmb::neumannForceComputation on the B200 solver = one isl_assemble_neumann_rows call per invocation; the
    reference's host loop (one insertToRHS per surface element) is not taken"""
    import re
    exe, args = RA.prepare(name, str(tmp_path))
    p = subprocess.run([os.path.join(APPS_B200, exe + "_mock")] + args, cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, ISL_MOCK_TRACE="1"))
    assert p.returncode == 0, p.stderr[-2000:]
    m = re.search(r"insert_lhs (\d+)\s+insert_rhs (\d+)\s+solve_cg (\d+)\s+assemble_neumann (\d+)", p.stderr)
    assert m, p.stderr[-500:]
    n_ins_lhs, n_ins_rhs, _, n_neumann = map(int, m.groups())
    assert n_neumann >= 1 and n_ins_lhs == 0 and n_ins_rhs == 0, m.group(0)


@needs_ref
def test_binding_included_too_late_is_a_compile_error(tmp_path):
    """without `-include insilico_b200_reference.hpp` the reference's driver facade would select the reference's generic
    host element loop for the B200 solver: the guard turns that into a compile error"""
    src = os.path.join(REFERENCE, "reference", "06-elastic", "compressibleWithDriver.cpp")
    cmd = ["/usr/bin/g++", "-std=c++17", "-DNDEBUG", "-w", "-fsyntax-only", "-DSPACEDIM=2",
           "-I" + os.path.join(ROOT, "tests", "ref_apps"), "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "oracle", "compat"), "-I" + REFERENCE, "-I" + os.path.dirname(src), src]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode != 0 and "no CPU fallback" in p.stderr
    p = subprocess.run(cmd[:-1] + ["-include", "insilico_b200_reference.hpp", src], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]


@needs_ref
def test_reference_app_mixed_poisson_golden(tmp_path):
    """reference/05-mixedPoisson (appTest.mk): stdout equals `ref`, the VTK output equals `ref.vtk`"""
    d = os.path.join(REFERENCE, "reference", "05-mixedPoisson")
    out = _run(str(tmp_path), "mixedPoisson", ["square_020.smf"], [os.path.join(d, "square_020.smf")])
    assert RA.same_output(out, open(os.path.join(d, "ref")).read(), rel=1e-6, noise=0.0)
    mine = open(os.path.join(str(tmp_path), "square_020.vtk")).read()
    assert RA.same_output(mine, open(os.path.join(d, "ref.vtk")).read(), rel=2e-5, noise=1e-9)


@needs_ref
def test_unsupported_kernel_is_a_compile_error(tmp_path):
    """reference/04-heat/convection.cpp uses heat::Convection, which the engine does not implement: no CPU fallback"""
    src = os.path.join(REFERENCE, "reference", "04-heat", "convection.cpp")
    cmd = ["/usr/bin/g++", "-std=c++17", "-DNDEBUG", "-w", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "ref_apps"),
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "compat"), "-I" + REFERENCE,
           "-include", "insilico_b200_reference.hpp", src]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode != 0 and "has no implementation in the B200 assembly engine" in p.stderr


@pytest.mark.skipif(not os.path.isdir(APPS_B200), reason="oracle/_ref/apps_b200 not built (needs /root/reference)")
def test_rescan_once_per_solver_policy_gives_the_same_newton_history(tmp_path):
    """b200_detail::rescanOncePerSolver(): the reference's objects are scanned at the first assembly call of each solver
    only (the reference's Newton loops change DoFs between solver instances)"""
    exe, args = RA.prepare("compressible_quad010", str(tmp_path))
    p = subprocess.run([os.path.join(APPS_B200, exe + "_mock")] + args, cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, ISL_RESCAN_PER_SOLVER="1"))
    assert p.returncode == 0, p.stderr[-2000:]
    expected = open(os.path.join(ROOT, "tests", "golden", "refrun_apps", "compressible_quad010.out")).read()
    assert RA.same_output(p.stdout, expected)
    # the reference's own semantics (every assembly call re-reads the objects) gives the same history as well; the
    # default is in between: full scan per solver + sampled comparison at the other calls
    p = subprocess.run([os.path.join(APPS_B200, exe + "_mock")] + args, cwd=str(tmp_path), capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, ISL_RESCAN_EVERY_CALL="1"))
    assert p.returncode == 0 and RA.same_output(p.stdout, expected), p.stderr[-2000:]
