#!/bin/bash
# Round 2, GPU call 5: full GPU suite with the new defaults, default bench line, launch list of the perturbed path
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== GPU test suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "== default bench"; timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; tail -c 3000 $O/bench_default.json; tail -5 $O/bench_default.err
echo "== launch list (structured then perturbed)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_prof_q1.csv python tools/prof_q1.py --steps 2 > $O/ncu_launch.log 2>&1; tail -3 $O/ncu_launch.log
grep -E "k_q1hex|k_check" $O/launches_prof_q1.csv | tail -12
echo "== ncu full of the two general kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_q1hex_elemK|k_q1hex_rows_fromK" -s 6 -c 2 -o $O/prof_fromk python tools/prof_q1.py --steps 2 --perturb-first 2>&1 | tail -3
} > $O/session5.log 2>&1
tail -90 $O/session5.log
