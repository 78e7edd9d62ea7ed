"""One GPU, BASELINE config 3 (Q2 hexahedra x 3 DoFs, St.Venant tangent) at a size whose matrix has more than 2^31 non-zeros
(n = 80: 512 000 elements, 2.4e9 non-zeros): 64-bit element -> CSR position maps.  The assembled entries are checked
against the oracle on sampled sub-meshes (tests/subbox_check.py).

    python tools/check_c3_large.py 80
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insilico_b200 import engine as E  # noqa: E402
from insilico_b200 import workloads  # noqa: E402
from tests import subbox_check as SB  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
    t0 = time.perf_counter()
    w = workloads.build("C3", n)
    t_build = time.perf_counter() - t0
    eng = E.Engine(0)
    w.upload(eng)
    eng.new_solver(w.n_eqn)
    t0 = time.perf_counter()
    w.register(eng)
    eng.synchronize()
    t_reg = time.perf_counter() - t0
    stream = torch.cuda.ExternalStream(eng.stream, device=0)
    ms = []
    for k in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.new_solver(w.n_eqn)
        a.record(stream)
        for op in w.ops:
            eng.stiffness_matrix_computation(op[1], op[2], op[3], op[4], op[5], incremental=op[6])
        eng.flush()
        b.record(stream)
        eng.synchronize(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    n_eq, nnz = eng.finish_assembly()
    print("n", n, "elements", len(w.conn), "equations", n_eq, "nnz", nnz, "(2^31 =", 1 << 31, ") host build s", round(t_build, 1),
          "register s", round(t_reg, 2), "ms per assembly", [round(x, 1) for x in ms], "free GB", round(torch.cuda.mem_get_info()[0] / 1e9, 1), flush=True)
    rp, col, val, rhs = eng.get_csr()
    eng.close()
    err, compared = SB.check_subboxes(w, rp, col, val, n_boxes=3, half_width=1.6 / n)
    print("C3_LARGE OK entries compared", compared, "worst", err)


if __name__ == "__main__":
    main()
