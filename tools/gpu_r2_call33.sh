#!/bin/bash
# Round 2, GPU call 33 (1 GPU): ncu --set full of k_tangent_hypel_sym<6> in its element-matrix-to-memory mode (C3 at 32^3)
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hypel_sym" -s 3 -c 1 -o $O/prof_hypel_sym_gather python bench.py --config C3 --size 32 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -2 | cut -c1-200
} > $O/session33.log 2>&1
tail -5 $O/session33.log
