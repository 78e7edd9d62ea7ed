#!/bin/bash
# Round 2, GPU call 42 (1 GPU): final state with the atomic-free generic path and the strip kernel as defaults: whole GPU
# suite, smoke, bench lines of every config, reference arm, ncu --set full of one C5 step
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== generic kernels: atomic / gathered scatter x per-entry / strips (gather reads the K row table)"
timeout 120 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "atomic_free_scatter" 2>&1 | tail -3 | tee $O/quick42.txt
if ! grep -q "52 passed" $O/quick42.txt; then echo "!! K row table fails: ISL_GEN_KROW=0 for the rest of the session"; export ISL_GEN_KROW=0; fi
echo "== whole GPU suite"
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -5
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1 | cut -c1-200
line() {
c=$1
python - <<PY
import json
l = json.load(open("$O/bench42_$c.json")); r = l["roofline"]
at = r.get("atomics")
print("$c", "ms", round(l["ms_per_step"], 3), "value %.4g" % l["value"], "frac", round(r["frac"], 4), "per op", [round(o["ms"], 2) for o in r.get("per_op_ms", [])], "atomics", (round(at["frac"], 3) if at else None), "cpu", l.get("cpu_baseline", {}).get("value"), "launches", l.get("gpu_launches"))
PY
}
timeout 200 python bench.py --config C5 > $O/bench42_C5.json 2> $O/bench42_C5.err; line C5
timeout 200 python bench.py > $O/bench42_default.json 2> $O/bench42_default.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/r2/bench42_default.json"))
print("C2 ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "nonaffine", l.get("roofline_nonaffine", {}).get("frac"), "e2e", l["e2e"]["ms_per_step"], "launches", l["gpu_launches"])
PY
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 | head -c 250; echo
for c in C4 C3 C1; do timeout 200 python bench.py --config $c > $O/bench42_$c.json 2> $O/bench42_$c.err; line $c; done
echo "== ncu --set full, one C5 step"
timeout 100 ncu --set full --clock-control none --import-source on -k regex:"k_gen_gather_rows|k_tangent" -s 6 -c 6 -f -o $O/prof42_C5 python bench.py --config C5 --no-e2e --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
ls -la $O/prof42_C5.ncu-rep
} > $O/session42.log 2>&1
tail -40 $O/session42.log
