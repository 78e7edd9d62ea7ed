#!/bin/bash
# Round 2, GPU call 35 (1 GPU): pipelined row gather: whole GPU suite, C3 / C4 / C5 / default bench lines, split of the two kernels
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== whole GPU suite"
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -5
for c in C3 C4 C5; do timeout 900 python bench.py --config $c > $O/bench35_$c.json 2> $O/bench35_$c.err; python - <<PY
import json
l = json.load(open("$O/bench35_$c.json")); r = l["roofline"]
print("$c", "ms", round(l["ms_per_step"], 3), "value %.4g" % l["value"], "frac", round(r["frac"], 4), "per op", [round(o["ms"], 2) for o in r["per_op_ms"]], "atomics", round(r["atomics"]["frac"], 3), "cpu", l.get("cpu_baseline", {}).get("value"))
PY
done
timeout 900 python bench.py > $O/bench35_default.json 2> $O/bench35_default.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/r2/bench35_default.json"))
print("C2 ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "nonaffine", l.get("roofline_nonaffine", {}).get("frac"), "e2e", l["e2e"]["ms_per_step"])
PY
echo "== C3 kernel split"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"hypel_sym|gather_rows" -s 4 -c 2 --csv --log-file $O/gather2_C3.csv python bench.py --config C3 --no-e2e --no-cpu-baseline --steps 2 > /dev/null 2>&1
grep -E "hypel_sym|gather_rows" $O/gather2_C3.csv | awk -F'","' '{print substr($5,1,50), $(NF-2), $(NF-1), $NF}' | cut -c1-140
} > $O/session35.log 2>&1
tail -30 $O/session35.log
