#!/bin/bash
# Round 2, GPU call 17 (1 GPU): Stokes slab generator on the engine (one-GPU gloo ranks), the blocks beyond 2^31 slot entries, default bench
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== multi-GPU host logic on one GPU"
timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu -k "one_gpu" 2>&1 | tail -5
echo "== C5 blocks at n = 80 (one GPU)"
timeout 1200 python tools/check_c5_slab_large.py 80 2>&1 | tail -6
echo "== default bench"
timeout 900 python bench.py > $O/bench17_default.json 2> $O/bench17_default.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/r2/bench17_default.json"))
print("C2 ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "e2e", l["e2e"]["ms_per_step"])
PY
} > $O/session17.log 2>&1
tail -40 $O/session17.log
