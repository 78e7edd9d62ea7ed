#!/bin/bash
# Round 2, GPU call 1: un-gated experimental tests, row-gather sweep, other configs on generic kernels.
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "== 1. GPU test suite (no -x)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== 2. experimental: row-gather kernels, device CG, tiled tangent"; ISL_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_parity_gpu.py -q -k rowgather 2>&1 | tail -30
ISL_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zz_linear_constraints.py -q -m gpu -k "cg or tiled" 2>&1 | tail -30
echo "== 4. row-gather sweep in one process (structured, then perturbed mesh)"
timeout 700 python tools/sweep_rows.py --n 256 --steps 10 --patch-rows 192,256,320 --stretch 1,2 --threads 256,128 2>&1 | tail -50
timeout 500 python tools/sweep_rows.py --n 256 --steps 10 --perturbed --patch-rows 192,256 --stretch 1,2 --threads 256,128 2>&1 | tail -30
echo "== 5. other configs (generic kernels), then with the register-tiled hyperelastic tangent"; timeout 600 python tools/bench_configs.py 2>&1 | tail -12
ISL_TANGENT_TILED=1 timeout 300 python tools/bench_configs.py --case stvenant_q2_hex --n 24 2>&1 | tail -2
ISL_TANGENT_TILED=1 timeout 300 python tools/bench_configs.py --case neohooke_p2_tet --n 24 2>&1 | tail -2
} > $O/session1.log 2>&1
tail -150 $O/session1.log
