#!/bin/bash
# Round 2, GPU call 25 (1 GPU): device DoF generation against the host function; its time at the C5 size
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== device dof generation"
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "dof_generation" 2>&1 | tail -6
timeout 900 python - <<'PY'
import time, numpy as np
from insilico_b200 import engine as E, meshgen
for shape, n, deg in ((E.TET, 64, 2), (E.HEX, 64, 2)):
    if shape == E.TET:
        coords, conn = meshgen.unit_cube_tet(n, n, n)
    else:
        coords, conn, _ = meshgen.unit_cube_hex(n, n, n)
    conn = meshgen.permute_elements(conn)
    eng = E.Engine(0); eng.set_mesh(shape, 1, coords, conn)
    eng.dof_generate(deg)
    t0 = time.perf_counter(); ed, nobj = eng.dof_generate(deg); t1 = time.perf_counter()
    ref, nref = E.dof_generate(shape, 1, conn, deg); t2 = time.perf_counter()
    print("shape", shape, "elements", len(conn), "objects", nobj, "device (incl. copy to host) ms", round((t1 - t0) * 1e3, 1), "host ms", round((t2 - t1) * 1e3, 1), "equal", bool(nobj == nref and np.array_equal(ed, ref)))
    eng.close()
PY
} > $O/session25.log 2>&1
tail -20 $O/session25.log
