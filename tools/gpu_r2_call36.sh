#!/bin/bash
# Round 2, GPU call 36 (1 GPU): Q2 tile kernel compiled for three (four) CTAs per SM: parity, C3 with and without
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== parity"
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_zy_full_size.py -q -m gpu -k "stvenant or neohooke or sub_mesh" 2>&1 | tail -3
for o in 0 1; do
ISL_HYPEL_OCC3=$o timeout 600 python bench.py --config C3 --no-e2e --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('occ3', $o, 'C3 ms', round(l['ms_per_step'],2), 'frac', round(l['roofline']['frac'],4))"
done
timeout 600 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"hypel_sym" -s 4 -c 1 --csv --log-file $O/occ3_C3.csv python bench.py --config C3 --no-e2e --no-cpu-baseline --steps 2 > /dev/null 2>&1
grep -E "hypel_sym" $O/occ3_C3.csv | awk -F'","' '{print substr($5,1,60), $(NF-2), $(NF-1), $NF}' | cut -c1-160
} > $O/session36.log 2>&1
tail -12 $O/session36.log
