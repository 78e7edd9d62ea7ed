#!/bin/bash
# Round 2, GPU call 37 (1 GPU): tile height of the hyperelastic kernel for P2 tetrahedra (MC 5 / 3 / 2): parity, C4; C3 with the new launch heuristic
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
for mc in 3 2; do
echo "== parity, MC $mc"
ISL_HYPEL_MC=$mc timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_zz_linear_constraints.py -q -m gpu -k "stvenant or neohooke or newton or mass_p2" 2>&1 | tail -2
done
for mc in 5 3 2; do
ISL_VERBOSE=1 ISL_HYPEL_MC=$mc timeout 600 python bench.py --config C4 --no-e2e --no-cpu-baseline --steps 5 2> $O/bench37_$mc.err | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('MC', $mc, 'C4 ms', round(l['ms_per_step'],2), [round(o['ms'],2) for o in l['roofline']['per_op_ms']], 'frac', round(l['roofline']['frac'],4))"
grep "tile kernel" $O/bench37_$mc.err | tail -1
done
ISL_VERBOSE=1 timeout 600 python bench.py --config C3 --no-e2e --no-cpu-baseline --steps 5 2> $O/bench37_C3.err | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('C3 ms', round(l['ms_per_step'],2), 'frac', round(l['roofline']['frac'],4))"
grep "tile kernel" $O/bench37_C3.err | tail -1
} > $O/session37.log 2>&1
tail -16 $O/session37.log
