#!/bin/bash
# Round 2, GPU call 29 (1 GPU): node-block position maps (fixed row length), all kernels: parity, C3 / C4 / C5 with and without
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== parity with block maps"
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_zz_linear_constraints.py tests/test_reference_run.py -q -m gpu 2>&1 | tail -5
for b in 0 1; do for c in C3 C4 C5; do
ISL_BLOCK_SLOTS=$b timeout 600 python bench.py --config $c --no-e2e --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('block_slots', $b, '$c', 'ms', round(l['ms_per_step'],2), [round(o['ms'],2) for o in l['roofline']['per_op_ms']], 'frac', round(l['roofline']['frac'],4))"
done; done
} > $O/session29.log 2>&1
tail -20 $O/session29.log
