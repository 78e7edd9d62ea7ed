#!/bin/bash
# Round 2, GPU call 2: ncu --set full of the row-gather kernels (affine + general) and timing of the alternatives
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== fp64 peak"; python -c "
from insilico_b200 import engine as E
e=E.Engine(0); print('fp64 TFLOP/s', e.measure_fp64_peak(), e.measure_fp64_peak())"
echo "== rows 256/128 structured + perturbed(update_coords)"
ISL_Q1_ROWS=1 ISL_PATCH_ROWS=256 ISL_ROWS_THREADS=128 python tools/prof_q1.py
echo "== rows general 256 thr"
ISL_Q1_ROWS=1 ISL_PATCH_ROWS=256 ISL_ROWS_THREADS=256 python tools/prof_q1.py
echo "== default patch kernels"
python tools/prof_q1.py
echo "== ncu full: rows kernels"
ISL_Q1_ROWS=1 ISL_PATCH_ROWS=256 ISL_ROWS_THREADS=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_q1hex_rows -s 4 -c 1 -o $O/prof_rows_affine python tools/prof_q1.py --steps 2 2>&1 | tail -5
ISL_Q1_ROWS=1 ISL_PATCH_ROWS=256 ISL_ROWS_THREADS=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_q1hex_rows_general -s 4 -c 1 -o $O/prof_rows_general python tools/prof_q1.py --steps 2 2>&1 | tail -5
} > $O/session2.log 2>&1
tail -60 $O/session2.log
