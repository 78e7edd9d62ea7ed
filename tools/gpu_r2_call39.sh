#!/bin/bash
# Round 2, GPU call 39 (1 GPU): atomic-free path of the generic kernels (ISL_GEN_GATHER): parity both ways, the suites with it
# switched on, C5 with and without it, launch list of C5; one host thread on two logical devices through the binding
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== generic kernels, atomic and atomic-free scatter; two devices from one thread"
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_reference_run.py -q -m gpu -k "atomic_free_scatter or two_devices" 2>&1 | tail -8
for g in 1 0; do
ISL_VERBOSE=1 ISL_GEN_GATHER=$g timeout 300 python bench.py --config C5 --no-e2e --no-cpu-baseline --steps 5 2> $O/bench39_C5_g$g.err > $O/bench39_C5_g$g.json
python - <<PY
import json
l = json.load(open("$O/bench39_C5_g$g.json")); r = l["roofline"]
print("gen_gather $g C5 ms", round(l["ms_per_step"], 3), "per op", [round(o["ms"], 2) for o in r["per_op_ms"]], "frac", round(r["frac"], 4), "launches", l["gpu_launches"])
PY
grep "atomic-free generic" $O/bench39_C5_g$g.err | tail -3
done
echo "== launch list, C5 with the gather"
ISL_GEN_GATHER=1 timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches39_C5.csv python bench.py --config C5 --no-e2e --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2/launches39_C5.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows[-40:]:
    name = r[4].split("(")[0][:60]; t = float(r[-1].replace(",", ""))
    agg.setdefault(name, []).append(t)
for k, v in agg.items(): print("  %-60s n %2d  avg %.1f us" % (k, len(v), sum(v) / len(v) / 1e3))
PY
echo "== suites with ISL_GEN_GATHER=1"
ISL_GEN_GATHER=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_reference_run.py tests/test_multigpu.py tests/test_zz_linear_constraints.py -q -m gpu -x 2>&1 | tail -6
} > $O/session39.log 2>&1
tail -40 $O/session39.log
