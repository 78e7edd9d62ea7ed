#!/bin/bash
# Round 2, GPU call 12 (1 GPU): binder identity fix, Mass kernel parity, final profiles of the default path
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== tests"; timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_zz_linear_constraints.py tests/test_reference_run.py -m gpu -q 2>&1 | tail -8
echo "== launch list of the default bench"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_launch12.log 2>&1; tail -2 $O/ncu_launch12.log
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$O/launches_default_r2.csv")) if len(r)>10 and r[0].isdigit()]
c=collections.Counter(); t=collections.Counter()
for r in rows:
    name=r[4].split("(")[0][-60:]; c[name]+=1; t[name]+=float(r[-1])
for k,v in t.most_common(12): print("%-62s n=%4d total %.3f ms"%(k,c[k],v/1e6))
PY
echo "== ncu full of the affine kernel (device-formed patches)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_q1hex_rows_affine -s 4 -c 1 -o $O/prof_rows_affine_v3 python tools/prof_q1.py --steps 2 2>&1 | tail -3
} > $O/session12.log 2>&1
tail -50 $O/session12.log
