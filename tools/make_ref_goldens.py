#!/usr/bin/env python
"""Generate tests/golden/refrun/*.npz by running the UNMODIFIED reference (oracle/_ref/ref_driver, i.e. the headers
under /root/reference compiled against the std-only Boost/Eigen stand-ins of oracle/compat) on the parity-test cases.

Only runs where /root/reference exists (this container); the fixtures it writes are committed and are what the
`-m "not gpu"` suite pins the oracle against and the `-m gpu` suite compares the CUDA engine with.

    make -C oracle ref && python tools/make_ref_goldens.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from insilico_b200 import engine as E  # noqa: E402  (host-side DoF logic only; no GPU needed)
from tests import flows  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
OUT = os.path.join(ROOT, "tests", "golden", "refrun")

SHAPE_NAME = {E.TRI: "triangle", E.QUAD: "quadrilateral", E.TET: "tetrahedron", E.HEX: "hexahedron"}
KERNEL_NAME = {E.K_LAPLACE: "laplace", E.K_VECTOR_LAPLACE: "vector_laplace", E.K_HYPEL_STVENANT: "stvenant",
               E.K_HYPEL_NEOHOOKE: "neohooke", E.K_PRESSURE_GRADIENT: "pressure_gradient",
               E.K_VELOCITY_DIVERGENCE: "velocity_divergence", E.K_MASS: "mass", E.K_CONVECTION: "convection"}

# (golden name, flows case, driver type, n, perturb, permute, register)
CASES = [
    ("laplace_q1_hex_n6", "laplace_q1_hex", "laplace_q1_hex", 6, True, False, False),
    ("laplace_q1_hex_n5_structured_reg", "laplace_q1_hex", "laplace_q1_hex", 5, False, False, True),
    ("laplace_q1_hex_n5_permuted", "laplace_q1_hex", "laplace_q1_hex", 5, True, True, False),
    ("laplace_q1_hex_values_n4", "laplace_q1_hex_values", "laplace_q1_hex", 4, True, False, False),
    ("laplace_q2_hex_n3", "laplace_q2_hex", "laplace_q2_hex", 3, True, False, False),
    ("laplace_p1_tet_n5", "laplace_p1_tet", "laplace_p1_tet", 5, True, False, False),
    ("laplace_q1_quad_n5", "laplace_q1_quad", "laplace_q1_quad", 5, True, False, False),
    ("laplace_p2_tri_n4", "laplace_p2_tri", "laplace_p2_tri", 4, True, False, False),
    ("vector_laplace_q1_hex_n4", "vector_laplace_q1_hex", "vector_laplace_q1_hex", 4, True, False, False),
    ("stvenant_q1_hex_n4", "stvenant_q1_hex", "solid_q1_hex", 4, True, False, False),
    ("neohooke_q1_hex_n4", "neohooke_q1_hex", "solid_q1_hex", 4, True, False, False),
    ("stvenant_q2_hex_n2", "stvenant_q2_hex", "solid_q2_hex", 2, True, False, False),
    ("stvenant_q1_quad_n5", "stvenant_q1_quad", "solid_q1_quad", 5, True, False, False),
    ("neohooke_p2_tet_n3", "neohooke_p2_tet", "solid_p2_tet", 3, True, False, False),
    ("stokes_p2p1_tet_n3", "stokes_p2p1_tet", "stokes_p2p1_tet", 3, True, False, True),
    ("stokes_q2q1_hex_n2", "stokes_q2q1_hex", "stokes_q2q1_hex", 2, True, False, False),
    ("stokes_q2q1_quad_n4", "stokes_q2q1_quad", "stokes_q2q1_quad", 4, True, False, True),
    ("stvenant_q2_quad_n4", "stvenant_q2_quad", "solid_q2_quad", 4, True, False, True),
    ("laplace_p1_tri_n6", "laplace_p1_tri", "laplace_p1_tri", 6, True, True, False),
    ("stvenant_p2_tet_n3", "stvenant_p2_tet", "solid_p2_tet", 3, True, False, False),
    # heat::Laplace with a conductivity function kappa(x) (setConductivityFunction, heat/Laplace.hpp:85-126)
    ("laplace_q1_hex_kappafun_n4", "laplace_q1_hex_kappafun", "laplace_q1_hex", 4, True, False, False),
    ("laplace_p2_tet_kappafun_n3", "laplace_p2_tet_kappafun", "laplace_p2_tet", 3, True, True, False),
    # base::kernel::Mass next to a stiffness matrix (the system of an implicit time step / a reaction term)
    ("mass_q1_hex_n4", "mass_q1_hex", "laplace_q1_hex", 4, True, False, False),
    ("mass_p2_tet_vector_n3", "mass_p2_tet_vector", "solid_p2_tet", 3, True, True, False),
    # general body forces f(x) (the caller's function evaluated per quadrature point, BodyForce.hpp:172-205)
    ("laplace_q1_hex_bodyfun_n4", "laplace_q1_hex_bodyfun", "laplace_q1_hex", 4, True, False, False),
    ("laplace_p2_tri_bodyfun_n4", "laplace_p2_tri_bodyfun", "laplace_p2_tri", 4, True, False, False),
    ("vector_laplace_q1_hex_bodyfun_n3", "vector_laplace_q1_hex_bodyfun", "vector_laplace_q1_hex", 3, True, True, False),
    # fluid::Convection on the tuple (u, u, u) (fluid/Convection.hpp:88-220)
    ("convection_q1_hex_n3", "convection_q1_hex", "vector_laplace_q1_hex", 3, True, False, False),
    ("convection_q2_quad_n4", "convection_q2_quad", "vector_laplace_q2_quad", 4, True, True, False),
    ("convection_q1_hex_at_rest_n3", "convection_q1_hex_at_rest", "vector_laplace_q1_hex", 3, True, False, False),
    # surface (Neumann) terms: base::asmb::neumannForceComputation over boundary meshes (NeumannForce.hpp, generateBoundaryMesh.hpp)
    ("neumann_q1_hex_n4", "neumann_q1_hex", "laplace_q1_hex", 4, True, False, False),
    ("neumann_p2_tet_solid_n2", "neumann_p2_tet_solid", "solid_p2_tet", 2, True, True, False),
    ("neumann_p2_tri_n4", "neumann_p2_tri", "laplace_p2_tri", 4, True, False, False),
    ("neumann_q1_quad_solid_n5", "neumann_q1_quad_solid", "solid_q1_quad", 5, True, False, True),
    # general linear constraints (slave DoFs with weighted ACTIVE masters, asmb/assembleMatrix.hpp:212-338)
    ("laplace_q1_hex_linear_n5", "laplace_q1_hex_linear", "laplace_q1_hex", 5, True, False, False),
    ("laplace_q1_hex_linear_n4_reg", "laplace_q1_hex_linear", "laplace_q1_hex", 4, False, False, True),
    ("laplace_q2_hex_linear_n2", "laplace_q2_hex_linear", "laplace_q2_hex", 2, True, False, False),
    ("laplace_p1_tet_linear_n4", "laplace_p1_tet_linear", "laplace_p1_tet", 4, True, False, True),
    ("stvenant_q1_hex_linear_n4", "stvenant_q1_hex_linear", "solid_q1_hex", 4, True, False, False),
    ("stokes_p2p1_tet_linear_n2", "stokes_p2p1_tet_linear", "stokes_p2p1_tet", 2, True, False, True),
]


BODYFUN_NAME = {"laplace_q1_hex_bodyfun": "hex_scalar", "laplace_p2_tri_bodyfun": "tri_scalar",
                "vector_laplace_q1_hex_bodyfun": "hex_vector"}


def case_name(case):
    return getattr(case, "name", None)


def write_smf(path, shape, coords, conn):
    """SMF as base/io/smf/Reader.hpp:66-330 reads it (insilico_b200.smf.write: 17 significant digits, doubles round-trip)"""
    from insilico_b200 import smf
    smf.write(path, shape, coords, conn)


def run_reference(case, driver_type, register, workdir, repeat=1, dump=True):
    """case: tests.flows.Case.  Returns dict of the reference's outputs."""
    smf = os.path.join(workdir, "mesh.smf")
    write_smf(smf, case.shape, case.coords, case.conn)
    out = os.path.join(workdir, "out")
    lines = ["type %s" % driver_type, "mesh %s" % smf, "out %s" % out, "register %d" % int(register),
             "repeat %d" % repeat, "dump %d" % int(dump)]
    for i, f in enumerate(case.fields):
        boundary, pin = int(f["boundary"]), int(f["pin"])
        pf = os.path.join(workdir, "presc%d.bin" % i)
        vf = os.path.join(workdir, "values%d.bin" % i)
        np.ascontiguousarray(f["presc"], dtype=np.float64).tofile(pf)
        np.ascontiguousarray(f["values"], dtype=np.float64).tofile(vf)
        line = "field %d %d %d %s %s" % (i, boundary, pin, pf, vf)
        if boundary == 2:    # Dirichlet on a part of the boundary: the status table says which DoF components
            sf = os.path.join(workdir, "status%d.bin" % i)
            np.ascontiguousarray(f["status"] == E.CONSTRAINED, dtype=np.float64).tofile(sf)
            line += " " + sf
        lines.append(line)
        for obj, comp, rhs, masters in f["linear"]:
            lines.append("constraint %d %d %d %.17g %d %s" % (i, obj, comp, rhs, len(masters), " ".join(
                "%d %d %.17g" % m for m in masters)))
    for op in case.ops:
        if op[0] == "matrix":
            lines.append("op matrix %s %d %d %d %s" % (KERNEL_NAME[op[1]], op[4], op[5], int(op[6]),
                                                       " ".join("%.17g" % p for p in op[2])))
        elif op[0] == "matrixfun":
            lines.append("op matrixfun kappa1 %d %d %d" % (op[4], op[5], int(op[6])))
        elif op[0] == "residual":
            lines.append("op residual %s %d %d 1 %s" % (KERNEL_NAME[op[1]], op[4], op[5],
                                                        " ".join("%.17g" % p for p in op[2])))
        elif op[0] == "bodyfun":
            lines.append("op bodyfun %s %d %d 1" % (BODYFUN_NAME[case_name(case)], op[3], op[3]))
        elif op[0] == "neumann":   # op neumann <force name> test test <face filter> params
            lines.append("op neumann %s %d %d %d %s" % (op[1], op[3], op[3], int(op[4]), " ".join("%.17g" % p for p in op[5])))
        elif op[0] == "body":
            lines.append("op body body %d %d 1 %s" % (op[3], op[3], " ".join("%.17g" % p for p in op[1])))
    job = os.path.join(workdir, "job.txt")
    with open(job, "w") as f:
        f.write("\n".join(lines) + "\n")
    proc = subprocess.run([DRIVER, job], capture_output=True, text=True, timeout=1800)
    if proc.returncode != 0:
        raise RuntimeError("%s failed (exit %d):\n%s" % (DRIVER, proc.returncode, proc.stderr[-3000:]))
    log = proc.stdout
    res = dict(log=log)
    if not dump:
        return res
    for i, f in enumerate(case.fields):
        res["elem_dof%d" % i] = np.loadtxt(out + ".f%d.elemdof.txt" % i, dtype=np.int64, ndmin=2).astype(np.int32)
        d = np.loadtxt(out + ".f%d.dofs.txt" % i, dtype=np.int64, ndmin=2)
        assert np.array_equal(d[:, 0], np.arange(len(d)))
        res["status%d" % i] = d[:, 1::2].astype(np.uint8)
        res["eqn%d" % i] = d[:, 2::2].astype(np.int64)
    res.update(read_system(out))
    return res


def read_system(out):
    """the finished system the driver dumped (debugLHS / debugRHS) as CSR"""
    lhs = np.loadtxt(out + ".lhs.txt", ndmin=2)
    rhs = np.loadtxt(out + ".rhs.txt", ndmin=2)
    n = len(rhs)
    r, c, v = lhs[:, 0].astype(np.int64), lhs[:, 1].astype(np.int64), lhs[:, 2]
    order = np.lexsort((c, r))
    r, c, v = r[order], c[order], v[order]
    assert not np.any((r[1:] == r[:-1]) & (c[1:] == c[:-1])), "duplicates in the finished matrix"
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    return dict(rowptr=np.cumsum(rowptr), col=c.astype(np.int32), val=v, rhs=rhs[:, 1].copy())


def main():
    os.makedirs(OUT, exist_ok=True)
    from oracle import oracle as orc  # checker; used here only to report the agreement while generating
    worst = 0.0
    only = sys.argv[1] if len(sys.argv) > 1 else ""     # e.g. `make_ref_goldens.py neumann`: the cases whose name contains it
    for gname, cname, dtype_, n, perturb, permute, register in CASES:
        if only not in gname:
            continue
        case = flows.build_case(cname, n, perturb, permute)
        with tempfile.TemporaryDirectory() as wd:
            ref = run_reference(case, dtype_, register, wd)
        ok_num = all(np.array_equal(ref["elem_dof%d" % i], f["elem_dof"]) and
                     np.array_equal(ref["status%d" % i], f["status"]) and
                     np.array_equal(ref["eqn%d" % i][f["status"] == 0], f["eqn"][f["status"] == 0])
                     for i, f in enumerate(case.fields))
        o = case.run_oracle(register=register)
        cmp_ = flows.compare((ref["rowptr"], ref["col"], ref["val"], ref["rhs"]), o)
        print("%-36s numbering %s  pattern %s  nnz %7d  val %.2e  rhs %.2e" % (
            gname, "exact" if ok_num else "DIFFERS", "exact" if cmp_["pattern_equal"] else "DIFFERS", cmp_["nnz"],
            cmp_.get("val_diff", np.nan), cmp_.get("rhs_diff", np.nan)))
        worst = max(worst, cmp_.get("val_diff", 1.0), cmp_.get("rhs_diff", 1.0))
        ref.pop("log")
        np.savez_compressed(os.path.join(OUT, gname + ".npz"), case=cname, n=n, perturb=perturb, permute=permute,
                            register=register, **ref)
    print("worst oracle-vs-reference difference: %.3e" % worst)


if __name__ == "__main__":
    main()
