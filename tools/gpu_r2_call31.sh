#!/bin/bash
# Round 2, GPU call 31 (1 GPU): atomic-free hyperelastic path (element matrices to memory + row gather): parity, C3 / C4 with and without
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== parity"
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_zz_linear_constraints.py tests/test_reference_run.py tests/test_zy_full_size.py -q -m gpu -k "stvenant or neohooke or solid or hook or sub_mesh or compressible or linearElastic or newton or mass_p2" 2>&1 | tail -6
for g in 0 1; do for c in C3 C4; do
ISL_VERBOSE=1 ISL_HYPEL_GATHER=$g timeout 600 python bench.py --config $c --no-e2e --no-cpu-baseline --steps 5 2> $O/bench31_${c}_$g.err | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('gather', $g, '$c', 'ms', round(l['ms_per_step'],2), [round(o['ms'],2) for o in l['roofline']['per_op_ms']], 'frac', round(l['roofline']['frac'],4), 'register ms', round(l['config']['register_fields_ms']))"
grep "atomic-free" $O/bench31_${c}_$g.err | tail -1
done; done
} > $O/session31.log 2>&1
tail -20 $O/session31.log
