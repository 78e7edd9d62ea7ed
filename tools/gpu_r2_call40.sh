#!/bin/bash
# Round 2, GPU call 40 (1 GPU): row buffers over the column range of the block (ISL_GEN_RANGE), staging budget with the gather,
# ncu --set full of one C5 step with the gather
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== generic kernels, atomic and atomic-free scatter (column-range buffers on)"
timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "atomic_free_scatter or stokes or registered" 2>&1 | tail -4
run() {   # label, env...
label=$1; shift
env "$@" ISL_VERBOSE=1 ISL_GEN_GATHER=1 timeout 300 python bench.py --config C5 --no-e2e --no-cpu-baseline --steps 5 2> $O/bench40_$label.err > $O/bench40_$label.json
python - <<PY
import json
l = json.load(open("$O/bench40_$label.json")); r = l["roofline"]
print("$label C5 ms", round(l["ms_per_step"], 3), "per op", [round(o["ms"], 2) for o in r["per_op_ms"]], "frac", round(r["frac"], 4))
PY
}
run range1 ISL_GEN_RANGE=1
grep "widest" $O/bench40_range1.err | tail -3
run range0 ISL_GEN_RANGE=0
run stage24 ISL_STAGE_KB=24
run stage96 ISL_STAGE_KB=96
echo "== ncu --set full, one C5 step with the gather"
ISL_GEN_GATHER=1 timeout 280 ncu --set full --clock-control none --import-source on -k regex:"k_gen_gather_rows|k_tangent" -s 6 -c 6 -f -o $O/prof40_C5_gather python bench.py --config C5 --no-e2e --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
ls -la $O/prof40_C5_gather.ncu-rep
} > $O/session40.log 2>&1
tail -30 $O/session40.log
