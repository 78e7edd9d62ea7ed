#!/bin/bash
# Round 2, GPU call 41 (1 GPU): register strips in the generic tangent kernel (ISL_GEN_TILE): parity in all four combinations, C5
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== generic kernels: atomic / gathered scatter x per-entry / strips"
timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "atomic_free_scatter" 2>&1 | tail -4
run() {   # label, env...
label=$1; shift
env "$@" ISL_GEN_GATHER=1 timeout 300 python bench.py --config C5 --no-e2e --no-cpu-baseline --steps 5 2> $O/bench41_$label.err > $O/bench41_$label.json
python - <<PY
import json
l = json.load(open("$O/bench41_$label.json")); r = l["roofline"]
print("$label C5 ms", round(l["ms_per_step"], 3), "per op", [round(o["ms"], 2) for o in r["per_op_ms"]], "frac", round(r["frac"], 4))
PY
}
run tile1 ISL_GEN_TILE=1
run tile1_stage32 ISL_GEN_TILE=1 ISL_STAGE_KB=32
run tile0 ISL_GEN_TILE=0
} > $O/session41.log 2>&1
tail -30 $O/session41.log
