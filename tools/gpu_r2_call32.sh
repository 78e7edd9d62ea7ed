#!/bin/bash
# Round 2, GPU call 32 (1 GPU): split of the atomic-free hyperelastic path into its two kernels (launch list), C3 and C4
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
for c in C3 C4; do
echo "== $c launch list"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct --clock-control none -k regex:"hypel_sym|gather_rows" -s 4 -c 2 --csv --log-file $O/gather_$c.csv python bench.py --config $c --no-e2e --no-cpu-baseline --steps 2 > /dev/null 2>&1
grep -E "hypel_sym|gather_rows" $O/gather_$c.csv | awk -F'","' '{print substr($5,1,60), $(NF-2), $(NF-1), $NF}' | cut -c1-160
done
echo "== newton loop test"
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "newton_loop" 2>&1 | tail -2
} > $O/session32.log 2>&1
tail -30 $O/session32.log
