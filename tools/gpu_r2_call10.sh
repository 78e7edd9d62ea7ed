#!/bin/bash
# Round 2, GPU call 10 (1 GPU): pipelined two-kernel general path, reworked scatter of the symmetric tangent kernel, binding scan
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== tests"; timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_zz_linear_constraints.py -m gpu -q 2>&1 | tail -6
echo "== Q1: structured + perturbed, chunks 16"; python tools/prof_q1.py
for nc in 1 4 32 64; do echo "== chunks $nc"; ISL_FROMK_CHUNKS=$nc python tools/prof_q1.py --perturb-first; done
echo "== bench configs"
for c in C3 C4; do timeout 600 python bench.py --config $c --steps 5 --no-cpu-baseline --no-e2e > $O/bench10_$c.json 2> $O/bench10_$c.err; python - <<PY
import json
l=json.load(open("$O/bench10_$c.json"))
print("$c", "value %.4g"%l["value"], "ms %.3f"%l["ms_per_step"], "frac %.4f"%l["roofline"]["frac"], l["roofline"]["per_op_ms"], "reg_ms %.0f"%l["config"]["register_fields_ms"])
PY
tail -3 $O/bench10_$c.err; done
echo "== default bench"; timeout 900 python bench.py --steps 10 > $O/bench10_default.json 2> $O/bench10_default.err; python - <<PY
import json
l=json.load(open("$O/bench10_default.json"))
print("C2 ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "nonaffine", l["roofline_nonaffine"]["frac"], l["roofline_nonaffine"]["kernel_ms"])
print("e2e", l["e2e"]["ms_per_step"], "api", l.get("e2e_reference_api",{}).get("ms_per_step"), l.get("e2e_reference_api",{}).get("value"))
print("newton", l.get("e2e_newton"))
print("cpu", l.get("cpu_baseline",{}).get("value"))
PY
tail -3 $O/bench10_default.err
} > $O/session10.log 2>&1
tail -60 $O/session10.log
