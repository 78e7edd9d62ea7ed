#!/bin/bash
# Round 2, GPU call 9 (1 GPU): cooperative scatter of the symmetric tangent kernel, Z-curve element order, Newton loop tests, smoke
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== tests"; timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_zz_linear_constraints.py tests/test_reference_run.py -m gpu -q 2>&1 | tail -8
echo "== bench configs"
for c in C3 C4 C5; do timeout 600 python bench.py --config $c --steps 5 --no-cpu-baseline --no-e2e > $O/bench9_$c.json 2> $O/bench9_$c.err; python - <<PY
import json
l=json.load(open("$O/bench9_$c.json"))
print("$c", "value %.4g"%l["value"], "ms %.3f"%l["ms_per_step"], "frac %.4f"%l["roofline"]["frac"], l["roofline"]["per_op_ms"], "reg_ms %.0f"%l["config"]["register_fields_ms"])
PY
tail -3 $O/bench9_$c.err; done
ISL_ELEM_ORDER=0 timeout 300 python bench.py --config C4 --steps 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C4 no elem order ms', l['ms_per_step'])"
echo "== default bench (reference api leg with the parallel scan)"; timeout 900 python bench.py --steps 10 > $O/bench9_default.json 2> $O/bench9_default.err; python - <<PY
import json
l=json.load(open("$O/bench9_default.json"))
print("C2 ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "nonaffine", l["roofline_nonaffine"]["frac"], l["roofline_nonaffine"]["kernel_ms"], "fp64", l["roofline"]["fp64"]["frac"])
print("e2e", l["e2e"]["ms_per_step"], "api", l.get("e2e_reference_api"))
print("cpu", l.get("cpu_baseline",{}).get("value"))
PY
tail -3 $O/bench9_default.err
echo "== ncu of the C3 tangent kernel"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tangent_hypel_sym -s 3 -c 1 -o $O/prof_hypel_sym python bench.py --config C3 --size 32 --steps 2 --no-cpu-baseline --no-e2e 2>&1 | tail -2
} > $O/session9.log 2>&1
tail -70 $O/session9.log
