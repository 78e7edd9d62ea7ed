#!/bin/bash
# Round 2, GPU call 26 (1 GPU): final state: whole GPU suite, smoke, bench lines of all configs, timing of the device DoF generation per stage
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== whole GPU suite"
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -6
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
echo "== device dof generation, timed"
ISL_VERBOSE=1 timeout 600 python - <<'PY'
import time, numpy as np
from insilico_b200 import engine as E, meshgen
for shape, n, deg in ((E.HEX, 64, 2), (E.TET, 64, 2), (E.HEX, 48, 3)):
    coords, conn = (meshgen.unit_cube_tet(n, n, n) if shape == E.TET else meshgen.unit_cube_hex(n, n, n)[:2])
    conn = meshgen.permute_elements(conn)
    eng = E.Engine(0); eng.set_mesh(shape, 1, coords, conn)
    eng.dof_generate(deg)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); ed, nobj = eng.dof_generate(deg); ts.append((time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter(); ref, nref = E.dof_generate(shape, 1, conn, deg); th = (time.perf_counter() - t0) * 1e3
    print("shape", shape, "degree", deg, "elements", len(conn), "objects", nobj, "device ms", [round(t, 1) for t in ts], "host ms", round(th, 1), "equal", bool(nobj == nref and np.array_equal(ed, ref)))
    eng.close()
PY
echo "== final bench lines"
for c in C1 C3 C4 C5; do timeout 900 python bench.py --config $c > $O/bench26_$c.json 2> $O/bench26_$c.err; python - <<PY
import json
l = json.load(open("$O/bench26_$c.json")); r = l["roofline"]
print("$c", "ms", round(l["ms_per_step"], 3), "value %.4g" % l["value"], "frac", round(r["frac"], 4), "hbm", round(r["hbm"]["frac"], 4), "fp64", round(r["fp64"]["frac"], 4), "cpu", l.get("cpu_baseline", {}).get("value"))
PY
done
timeout 900 python bench.py > $O/bench26_default.json 2> $O/bench26_default.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/r2/bench26_default.json"))
print("C2 ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "nonaffine", l.get("roofline_nonaffine", {}).get("frac"), "e2e", l["e2e"]["ms_per_step"], "launches", l["gpu_launches"], "clocks", l["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | head -c 300; echo
} > $O/session26.log 2>&1
tail -40 $O/session26.log
