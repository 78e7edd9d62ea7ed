#!/usr/bin/env python
"""Numbering fixtures at larger sizes (numbering parity was pinned at n <= 6 only): the UNMODIFIED reference
(oracle/_ref/ref_driver, `dump 2` = DoF numbering only) numbers the DoFs of Q2 hexahedra, P2 tetrahedra and the
Taylor-Hood pair on meshes with 16 / 12 elements per direction, perturbed and in permuted element order; the SHA-256 of
its element -> DoF table, DoF status and equation numbers go to tests/golden/refrun/numbering_digests.json
(tests/test_host_logic.py compares the product's base/dof restatement and the oracle's with them).

    python tools/make_ref_numbering.py        # needs /root/reference (oracle/_ref built)
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import flows  # noqa: E402
from tools import make_ref_goldens as G  # noqa: E402

CASES = [("stvenant_q2_hex", "solid_q2_hex", 16, True, True), ("neohooke_p2_tet", "solid_p2_tet", 16, True, True),
         ("stokes_p2p1_tet", "stokes_p2p1_tet", 12, True, True), ("laplace_q2_hex", "laplace_q2_hex", 16, True, False)]


def digest(a, dtype):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=dtype).tobytes()).hexdigest()


def case_digests(elem_dofs, statuses, eqns):
    return {"fields": [{"elem_dof": digest(ed, np.int32), "status": digest(st, np.uint8), "eqn": digest(eq, np.int64),
                        "n_obj": int(len(st)), "n_active": int((np.asarray(eq) >= 0).sum())}
                       for ed, st, eq in zip(elem_dofs, statuses, eqns)]}


def main():
    out = {}
    for name, dtype_name, n, perturb, permute in CASES:
        c = flows.build_case(name, n, perturb, permute)
        with tempfile.TemporaryDirectory() as wd:
            smf = os.path.join(wd, "mesh.smf")
            G.write_smf(smf, c.shape, c.coords, c.conn)
            lines = ["type %s" % dtype_name, "mesh %s" % smf, "out %s/out" % wd, "register 0", "repeat 1", "dump 2"]
            for i, f in enumerate(c.fields):
                pf, vf = os.path.join(wd, "p%d.bin" % i), os.path.join(wd, "v%d.bin" % i)
                np.ascontiguousarray(f["presc"], dtype=np.float64).tofile(pf)
                np.ascontiguousarray(f["values"], dtype=np.float64).tofile(vf)
                lines.append("field %d %d %d %s %s" % (i, int(f["boundary"]), int(f["pin"]), pf, vf))
            job = os.path.join(wd, "job.txt")
            open(job, "w").write("\n".join(lines) + "\n")
            import subprocess
            subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_driver"), job], check=True, capture_output=True, timeout=1800)
            eds, sts, eqs = [], [], []
            for i in range(len(c.fields)):
                eds.append(np.loadtxt("%s/out.f%d.elemdof.txt" % (wd, i), dtype=np.int64, ndmin=2))
                d = np.loadtxt("%s/out.f%d.dofs.txt" % (wd, i), dtype=np.int64, ndmin=2)
                sts.append(d[:, 1::2]); eqs.append(d[:, 2::2])
        key = "%s_n%d" % (name, n)
        out[key] = case_digests(eds, sts, eqs)
        out[key].update(perturb=perturb, permute=permute, n_elems=int(len(c.conn)))
        print(key, out[key]["fields"][0]["n_obj"], out[key]["fields"][0]["n_active"])
    with open(os.path.join(ROOT, "tests", "golden", "refrun", "numbering_digests.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
