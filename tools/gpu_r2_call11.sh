#!/bin/bash
# Round 2, GPU call 11 (1 GPU): patch formation on the device, sub-mesh parity at config sizes, full suite
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== Q1 default (device patches) / host patches"; ISL_VERBOSE=1 python tools/prof_q1.py 2>&1 | grep -v "batch slot"
ISL_PATCH_HOST=1 python tools/prof_q1.py
for r in 192 320; do echo "== rows $r"; ISL_PATCH_ROWS=$r python tools/prof_q1.py; done
echo "== stretch 2 / 4"; ISL_PATCH_STRETCH=2 python tools/prof_q1.py; ISL_PATCH_STRETCH=4 python tools/prof_q1.py
echo "== GPU test suite"; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12
echo "== default bench"; timeout 900 python bench.py --steps 10 > $O/bench11_default.json 2> $O/bench11_default.err; python - <<PY
import json
l=json.load(open("$O/bench11_default.json"))
print("C2 ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "register_ms", l["config"]["register_fields_ms"], "nonaffine", l["roofline_nonaffine"]["frac"])
print("e2e", l["e2e"]["ms_per_step"], "api", l.get("e2e_reference_api",{}).get("ms_per_step"), l.get("e2e_reference_api",{}).get("value"), "cpu", l.get("cpu_baseline",{}).get("value"))
PY
tail -3 $O/bench11_default.err
} > $O/session11.log 2>&1
tail -60 $O/session11.log
