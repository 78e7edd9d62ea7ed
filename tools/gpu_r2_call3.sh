#!/bin/bash
# Round 2, GPU call 3: stencil-sum row kernel with bulk-copy write-out: parity, sweep, ncu
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== parity"; timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "rowgather" 2>&1 | tail -15
echo "== rows 256/128 structured + perturbed(update_coords)"
ISL_Q1_ROWS=1 ISL_PATCH_ROWS=256 ISL_ROWS_THREADS=128 python tools/prof_q1.py
echo "== sweep"
timeout 600 python tools/sweep_rows.py --n 256 --steps 10 --patch-rows 192,256,320 --stretch 1,2,3 --threads 128,256 2>&1 | tail -25
echo "== ncu full"
ISL_Q1_ROWS=1 ISL_PATCH_ROWS=256 ISL_ROWS_THREADS=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_q1hex_rows_affine -s 4 -c 1 -o $O/prof_rows_affine_v2 python tools/prof_q1.py --steps 2 2>&1 | tail -4
} > $O/session3.log 2>&1
tail -70 $O/session3.log
