#!/bin/bash
# Round 2, GPU call 23 (1 GPU): 64-bit slot maps (forced at test sizes, then C3 at 80^3 = 2.4e9 non-zeros), staging-budget sweep of the generic kernels
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== parity with 64-bit slot maps forced"
ISL_SLOT64=1 timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_zz_linear_constraints.py -q -m gpu -x 2>&1 | tail -4
echo "== staging budget sweep (C5, C4, C3 at 32^3)"
for kb in 96 64 48 32 24; do
  for c in C5 C4; do
    ISL_STAGE_KB=$kb timeout 600 python bench.py --config $c --no-e2e --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('stage_kb', $kb, '$c', 'ms', round(l['ms_per_step'],2), [round(o['ms'],2) for o in l['roofline']['per_op_ms']])"
  done
done
echo "== C3 at 80^3 on one GPU (more than 2^31 non-zeros), sub-mesh oracle parity"
timeout 1500 python tools/check_c3_large.py 80 2>&1 | tail -5
} > $O/session23.log 2>&1
tail -40 $O/session23.log
