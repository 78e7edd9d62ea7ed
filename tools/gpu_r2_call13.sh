#!/bin/bash
# Round 2, GPU call 13 (1 GPU): producer/consumer pipeline kernel for general Q1 elements; Mass kernel failure details
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== mass failure"; timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "mass_q1_hex-7" 2>&1 | tail -30
echo "== pipeline parity"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "rowgather or laplace_q1_hex" 2>&1 | tail -6
echo "== Q1: pipeline"; ISL_VERBOSE=1 timeout 300 python tools/prof_q1.py 2>&1 | grep -v "batch slot"
for pr in 32768 131072; do echo "== pipe rows $pr"; ISL_PIPE_ROWS=$pr timeout 300 python tools/prof_q1.py --perturb-first; done
echo "== two kernels"; ISL_FROMK_PIPELINE=0 timeout 300 python tools/prof_q1.py --perturb-first
echo "== full-size perturbed parity"; timeout 900 python -m pytest tests/test_zy_full_size.py -m gpu -q -k "perturbed" 2>&1 | tail -4
} > $O/session13.log 2>&1
tail -70 $O/session13.log
