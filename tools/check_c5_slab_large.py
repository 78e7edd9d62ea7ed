"""One GPU, one slab = the whole cube of 6 n^3 tetrahedra from partition.structured_stokes_slab: the Stokes blocks at a
size where element * (rows x columns) exceeds 2^31 (n >= 74), checked against the oracle on sampled sub-meshes.

    python tools/check_c5_slab_large.py 80
"""
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insilico_b200 import engine as E  # noqa: E402
from insilico_b200 import partition  # noqa: E402
from tests import subbox_check as SB  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
    t0 = time.perf_counter()
    wl = partition.structured_stokes_slab(n, 0, 1)
    t_gen = time.perf_counter() - t0
    ops = [("matrix", E.K_VECTOR_LAPLACE, [1.0], 4, 0, 0, True), ("matrix", E.K_PRESSURE_GRADIENT, [0.0], 4, 0, 1, True),
           ("matrix", E.K_VELOCITY_DIVERGENCE, [0.0], 4, 1, 0, True)]
    eng = E.Engine(0)
    eng.set_mesh(E.TET, 1, wl["coords"], wl["conn"])
    for i, f in enumerate(wl["fields"]):
        eng.set_field(i, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"], f["eqn"], f["status"], f["presc"], f["values"])
    eng.new_solver(wl["n_eqn_local"])
    t0 = time.perf_counter()
    for op in ops:
        eng.register_fields(op[4], op[5])
    eng.synchronize()
    t_reg = time.perf_counter() - t0
    stream = torch.cuda.ExternalStream(eng.stream, device=0)
    ms = []
    for k in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.new_solver(wl["n_eqn_local"])
        a.record(stream)
        for op in ops:
            eng.stiffness_matrix_computation(op[1], op[2], op[3], op[4], op[5], incremental=op[6])
        eng.flush()
        b.record(stream)
        eng.synchronize(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    rp, col, val, rhs = eng.get_csr()
    print("n", n, "tets", len(wl["conn"]), "slot entries of UU", len(wl["conn"]) * 900, "nnz", len(col), "generate s", round(t_gen, 2),
          "register s", round(t_reg, 2), "ms per step", [round(x, 2) for x in ms], "free GB", torch.cuda.mem_get_info()[0] / 1e9)
    eng.close()
    w = SimpleNamespace(coords=wl["coords"], conn=wl["conn"], fields=wl["fields"], ops=ops, shape=E.TET, geom_deg=1, dim=3)
    err, compared = SB.check_subboxes(w, rp, col, val, n_boxes=4, half_width=2.4 / n)
    print("C5_SLAB_LARGE OK entries compared", compared, "worst", err)


if __name__ == "__main__":
    main()
