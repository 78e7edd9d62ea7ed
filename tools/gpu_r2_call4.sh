#!/bin/bash
# Round 2, GPU call 4: new defaults (row kernels) through the whole GPU suite; two-kernel general path; bench configs
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== default: structured + perturbed(update_coords)"; ISL_VERBOSE=1 python tools/prof_q1.py
echo "== perturbed first"; ISL_VERBOSE=1 python tools/prof_q1.py --perturb-first
echo "== GPU test suite"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== bench configs"
for c in C1 C3 C4 C5; do timeout 600 python bench.py --config $c --steps 5 > $O/bench_$c.json 2> $O/bench_$c.err; tail -c 2500 $O/bench_$c.json; tail -3 $O/bench_$c.err; done
} > $O/session4.log 2>&1
tail -80 $O/session4.log
