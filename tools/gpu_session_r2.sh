#!/bin/bash
# First GPU call of round 2 (round 1 ended without GPU minutes for the last session):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_session_r2.sh'
# With 2 GPUs (gpurun --gpus 2) also: ISL_TEST_EXPERIMENTAL=1 python -m pytest tests/test_multigpu.py -q   (general partition)
# Everything lands in gpurun_out/r2/.  Steps are independent: a failing one does not stop the others.
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== 1. GPU test suite"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40
echo "== 2. experimental: row-gather kernels, device CG"; ISL_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_parity_gpu.py -q -k rowgather 2>&1 | tail -8
ISL_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zz_linear_constraints.py -q -m gpu -k "cg or tiled" 2>&1 | tail -8
echo "== 3. default bench"; timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; tail -c 1500 $O/bench_default.json
echo "== 4. row-gather sweep in one process (structured, then perturbed mesh)"
timeout 900 python tools/sweep_rows.py --n 256 --steps 10 2>&1 | tail -90
timeout 600 python tools/sweep_rows.py --n 256 --steps 10 --perturbed --patch-rows 192,256 --stretch 1,2 --threads 256,128 2>&1 | tail -40
echo "== 5. other configs (generic kernels), then with the register-tiled hyperelastic tangent"; timeout 900 python tools/bench_configs.py 2>&1 | tail -12
ISL_TANGENT_TILED=1 timeout 600 python tools/bench_configs.py --case stvenant_q2_hex --n 24 2>&1 | tail -2
ISL_TANGENT_TILED=1 timeout 600 python tools/bench_configs.py --case neohooke_p2_tet --n 24 2>&1 | tail -2
echo "== 6. launch list of the default bench"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_launch.log 2>&1; tail -3 $O/ncu_launch.log
} > $O/session.log 2>&1
tail -60 $O/session.log
