#!/bin/bash
# Round 2, GPU call 14 (1 GPU): full suite (Mass, kappa(x), binder identity), row-kernel variants of the general path, bench
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== GPU test suite"; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12
for v in 0 1 2 3 4; do echo "== general path, row-kernel variant $v"; ISL_FROMK_VARIANT=$v timeout 300 python tools/prof_q1.py --perturb-first; done
echo "== default bench"; timeout 900 python bench.py --steps 10 > $O/bench14_default.json 2> $O/bench14_default.err; python - <<PY
import json
l=json.load(open("$O/bench14_default.json"))
print("C2 ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "register_ms", l["config"]["register_fields_ms"], "nonaffine", l["roofline_nonaffine"]["frac"], l["roofline_nonaffine"]["kernel_ms"])
print("e2e", l["e2e"]["ms_per_step"], "api", l.get("e2e_reference_api",{}).get("ms_per_step"), l.get("e2e_reference_api",{}).get("value"), "cpu", l.get("cpu_baseline",{}).get("value"))
print("newton", {k:v for k,v in l.get("e2e_newton",{}).items() if k!="what"})
PY
tail -3 $O/bench14_default.err
} > $O/session14.log 2>&1
tail -60 $O/session14.log
