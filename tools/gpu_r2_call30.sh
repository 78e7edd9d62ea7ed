#!/bin/bash
# Round 2, GPU call 30 (1 GPU): measured FP64 atomic-add rates and the bench lines of configs 3-5 with their atomic fraction
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== atomic-add rates"
timeout 300 python -c "
from insilico_b200 import engine as E
eng = E.Engine(0)
print('G atomic adds/s: coalesced', round(eng.measure_red_peak(0), 1), ' triples at random places', round(eng.measure_red_peak(1), 1), ' singles at random places', round(eng.measure_red_peak(2), 1))
eng.close()"
for c in C3 C4 C5; do timeout 900 python bench.py --config $c > $O/bench30_$c.json 2> $O/bench30_$c.err; python - <<PY
import json
l = json.load(open("$O/bench30_$c.json")); r = l["roofline"]
print("$c", "ms", round(l["ms_per_step"], 3), "value %.4g" % l["value"], "frac", round(r["frac"], 4), "atomics", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r["atomics"].items() if k != "peak_source"})
PY
done
} > $O/session30.log 2>&1
tail -12 $O/session30.log
