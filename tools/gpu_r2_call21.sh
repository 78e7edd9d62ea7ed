#!/bin/bash
# Round 2, GPU call 21 (1 GPU): launch list of the default bench command, ncu of the generic Stokes / force kernels (C5, C4),
# traffic of the headline kernel with device-formed patches, final bench lines of all configs
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== launch list of the default bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default_r2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_default.log 2>&1; tail -2 $O/ncu_default.log | cut -c1-300
grep -c "k_" $O/launches_default_r2.csv
echo "== ncu full: headline kernel (device-formed patches)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_q1hex_rows_affine" -s 4 -c 1 -o $O/prof_rows_affine_v4 python tools/prof_q1.py --steps 2 2>&1 | tail -2
echo "== ncu full: k_tangent on the Stokes blocks (C5 at 6*32^3) and k_force (C4)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tangent" -s 9 -c 3 -o $O/prof_c5_tangent python bench.py --config C5 --size 32 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -2 | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force" -s 3 -c 1 -o $O/prof_c4_force python bench.py --config C4 --size 32 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -2 | cut -c1-200
echo "== final bench lines"
for c in C1 C3 C4 C5; do timeout 900 python bench.py --config $c > $O/bench21_$c.json 2> $O/bench21_$c.err; python - <<PY
import json
l = json.load(open("$O/bench21_$c.json")); r = l["roofline"]
print("$c", "ms", round(l["ms_per_step"], 3), "value %.4g" % l["value"], "frac", round(r["frac"], 4), "hbm", round(r["hbm"]["frac"], 4), "fp64", round(r["fp64"]["frac"], 4), "cpu", l.get("cpu_baseline", {}).get("value"))
PY
done
timeout 900 python bench.py > $O/bench21_default.json 2> $O/bench21_default.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/r2/bench21_default.json"))
print("C2 ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "nonaffine", l.get("roofline_nonaffine", {}).get("frac"), "e2e", l["e2e"]["ms_per_step"], "api", l["config"].get("e2e_reference_api", l.get("e2e_reference_api")))
PY
} > $O/session21.log 2>&1
tail -40 $O/session21.log
