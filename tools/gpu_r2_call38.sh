#!/bin/bash
# Round 2, GPU call 38 (1 GPU): final state of the round: whole GPU suite, smoke, bench lines of every config, reference arm
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== whole GPU suite"
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1 | cut -c1-200
for c in C1 C3 C4 C5; do timeout 900 python bench.py --config $c > $O/bench38_$c.json 2> $O/bench38_$c.err; python - <<PY
import json
l = json.load(open("$O/bench38_$c.json")); r = l["roofline"]
at = r.get("atomics")
print("$c", "ms", round(l["ms_per_step"], 3), "value %.4g" % l["value"], "frac", round(r["frac"], 4), "per op", [round(o["ms"], 2) for o in r["per_op_ms"]], "atomics", (round(at["frac"], 3) if at else None), "cpu", l.get("cpu_baseline", {}).get("value"))
PY
done
timeout 900 python bench.py > $O/bench38_default.json 2> $O/bench38_default.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/r2/bench38_default.json"))
print("C2 ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "nonaffine", l.get("roofline_nonaffine", {}).get("frac"), "e2e", l["e2e"]["ms_per_step"], "launches", l["gpu_launches"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | head -c 250; echo
} > $O/session38.log 2>&1
tail -30 $O/session38.log
