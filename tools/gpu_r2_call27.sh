#!/bin/bash
# Round 2, GPU call 27 (1 GPU): general Q1 path with the row tiles launched along a Z-curve (L2 reuse of the element matrices)
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== general path, tile order off / on"
ISL_FROMK_TILE_ORDER=0 timeout 600 python tools/prof_q1.py --steps 10 --perturb-first 2>&1 | tail -1
ISL_FROMK_TILE_ORDER=1 timeout 600 python tools/prof_q1.py --steps 10 --perturb-first 2>&1 | tail -1
echo "== parity of the general path"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_zy_full_size.py -q -m gpu -k "laplace_q1_hex or mass_q1 or perturbed or rowgather or full" 2>&1 | tail -4
echo "== dram traffic of the row kernel with the tile order"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"k_q1hex_rows_fromK|k_q1hex_elemK" -s 6 -c 2 --csv --log-file $O/fromk_tile_order.csv python tools/prof_q1.py --steps 2 --perturb-first > /dev/null 2>&1
grep -E "k_q1hex" $O/fromk_tile_order.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | cut -c1-200
} > $O/session27.log 2>&1
tail -30 $O/session27.log
