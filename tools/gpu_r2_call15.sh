#!/bin/bash
# Round 2, GPU call 15 (1 GPU): asynchronous hand-off (pipelined e2e), faster SpMV of the device CG, default general variant
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== tests"; timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_zz_linear_constraints.py -m gpu -q 2>&1 | tail -6
echo "== default bench"; timeout 900 python bench.py --steps 10 > $O/bench15_default.json 2> $O/bench15_default.err; python - <<PY
import json
l=json.load(open("$O/bench15_default.json"))
print("C2 ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "register_ms", l["config"]["register_fields_ms"], "nonaffine", l["roofline_nonaffine"]["frac"], l["roofline_nonaffine"]["kernel_ms"])
print("e2e", l["e2e"]["ms_per_step"], "serial", l["e2e"]["ms_per_step_serial"], "api", l.get("e2e_reference_api",{}).get("ms_per_step"), l.get("e2e_reference_api",{}).get("value"), "cpu", l.get("cpu_baseline",{}).get("value"))
print("newton", {k:v for k,v in l.get("e2e_newton",{}).items() if k!="what"})
PY
tail -3 $O/bench15_default.err
} > $O/session15.log 2>&1
tail -30 $O/session15.log
