#!/bin/bash
# Round 2, GPU call 6 (1 GPU): parity of the symmetric tangent kernel + element order + fromK v2, timings
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== GPU test suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "== Q1 default: structured + perturbed"; python tools/prof_q1.py
echo "== bench configs"
for c in C3 C4 C5; do timeout 600 python bench.py --config $c --steps 5 --no-cpu-baseline > $O/bench6_$c.json 2> $O/bench6_$c.err; python - <<PY
import json
l=json.load(open("$O/bench6_$c.json"))
print("$c", "value %.4g"%l["value"], "ms %.3f"%l["ms_per_step"], "frac %.4f"%l["roofline"]["frac"], l["roofline"]["per_op_ms"], "reg_ms %.0f"%l["config"]["register_fields_ms"])
PY
tail -3 $O/bench6_$c.err; done
echo "== C3 without sym / without element order"
ISL_TANGENT_SYM=0 timeout 300 python bench.py --config C3 --n 32 --steps 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C3 n32 nosym ms', l['ms_per_step'])"
timeout 300 python bench.py --config C3 --n 32 --steps 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C3 n32 sym ms', l['ms_per_step'])"
ISL_ELEM_ORDER=0 timeout 300 python bench.py --config C5 --steps 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5 no elem order ms', l['ms_per_step'])"
} > $O/session6.log 2>&1
tail -60 $O/session6.log
