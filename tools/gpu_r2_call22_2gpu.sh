#!/bin/bash
# Round 2, GPU call 22 (2 GPUs): the driver's scaling command at N=2 with its default flags (e2e with the overlapped hand-off), weak and strong
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
echo "== N=2 default (weak, e2e on)"
timeout 900 $TR --nproc-per-node 2 --master-port 29621 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench22_n2.json 2> $O/bench22_n2.err
python - <<'PY'
import json
txt = open("gpurun_out/r2/bench22_n2.json").read()
print("stdout lines:", len(txt.strip().splitlines()))
l = json.loads(txt)
print("N=2 ms", l["ms_per_step"], "value %.4g" % l["value"], "e2e", l["e2e"])
PY
grep -E "rror|Traceback" $O/bench22_n2.err | tail -3
echo "== N=2 strong"
timeout 900 $TR --nproc-per-node 2 --master-port 29622 bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong --no-e2e > $O/bench22_n2_strong.json 2> $O/bench22_n2_strong.err
python -c "import json; l=json.load(open('$O/bench22_n2_strong.json')); print('N=2 strong ms', l['ms_per_step'])"
echo "== reference arm under torchrun"
timeout 900 $TR --nproc-per-node 2 --master-port 29623 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench22_ref.json 2> $O/bench22_ref.err; head -c 400 $O/bench22_ref.json; echo
echo "== two-GPU tests"
timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu 2>&1 | tail -3
} > $O/session22.log 2>&1
tail -30 $O/session22.log
