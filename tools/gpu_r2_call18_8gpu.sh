#!/bin/bash
# Round 2, GPU call 18 (8 GPUs): driven-cavity blocks (C5) on slabs generated per rank: parity at 8 ranks, then 24.6 M and 37.4 M tetrahedra
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
echo "== stokes slabs, 8 ranks, against the oracle"
timeout 300 $TR --nproc-per-node 8 --master-port 29611 tests/dist_check_general.py stokes_slab 16 2>&1 | grep -E "DIST_CHECK|rror" | tail -3
echo "== two-GPU tests the driver can run"
timeout 600 python -m pytest tests/test_multigpu.py -q -m gpu -k "two_gpu" 2>&1 | tail -3
for n in 160 184; do
echo "== C5 n=$n over 8 GPUs"
timeout 900 $TR --nproc-per-node 8 --master-port 296$((n % 90 + 10)) bench.py --gpus 8 --config C5 --size $n --steps 5 --no-e2e > $O/bench_n8_C5_$n.json 2> $O/bench_n8_C5_$n.err
python - <<PY
import json
try:
    l = json.load(open("$O/bench_n8_C5_$n.json"))
    print("n", $n, "tets", l["config"]["n_elems"], "ms", l["ms_per_step"], "value", l["value"], "nnz/gpu", l["config"]["nnz_per_gpu"], "register ms", l["config"]["register_fields_ms"], "per op", [round(o["ms"], 2) for o in l["roofline"]["per_op_ms"]])
except Exception as e:
    print("no line:", e)
PY
grep -E "rror|memory" $O/bench_n8_C5_$n.err | tail -4
done
} > $O/session18.log 2>&1
tail -40 $O/session18.log
