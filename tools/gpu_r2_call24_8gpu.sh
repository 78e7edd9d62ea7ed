#!/bin/bash
# Round 2, GPU call 24 (8 GPUs): BASELINE config 5 at its named size: 50 M tetrahedra over 8 GPUs (n = 203: 50.2 M), 64-bit slot maps
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
for n in 203; do
echo "== C5 n=$n over 8 GPUs"
timeout 1000 $TR --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --config C5 --size $n --steps 4 --no-e2e > $O/bench_n8_C5_$n.json 2> $O/bench_n8_C5_$n.err
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
} > $O/session24.log 2>&1
tail -20 $O/session24.log
