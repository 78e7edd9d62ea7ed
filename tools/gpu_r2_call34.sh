#!/bin/bash
# Round 2, GPU call 34 (1 GPU): hyperelastic element kernel with the displacement gradient formed by all threads: parity, C3 / C4
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== parity"
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_zz_linear_constraints.py tests/test_reference_run.py -q -m gpu -k "stvenant or neohooke or solid or hook or compressible or linearElastic or newton or mass_p2" 2>&1 | tail -4
for c in C3 C4; do
timeout 600 python bench.py --config $c --no-e2e --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('$c', 'ms', round(l['ms_per_step'],2), [round(o['ms'],2) for o in l['roofline']['per_op_ms']], 'frac', round(l['roofline']['frac'],4))"
done
} > $O/session34.log 2>&1
tail -12 $O/session34.log
