#!/bin/bash
# Round 2, GPU call 8 (2 GPUs): interface patches on the communication stream; C5 over 2 GPUs; device Newton loop tests
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_multigpu.py -q -k "two_gpu" 2>&1 | tail -5
echo "== newton / cg tests"; timeout 600 python -m pytest tests/test_zz_linear_constraints.py -q -m gpu -k "newton or cg" 2>&1 | tail -8
for rep in 1 2; do
echo "== N=2 weak (rep $rep)"; BENCH_ALL_RANKS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$rep bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > $O/bench_n2_weak.json 2> $O/bench_n2_weak.err; python -c "import json; l=json.load(open('$O/bench_n2_weak.json')); print('N=2 weak ms', l['ms_per_step'], 'value', l['value'])"; grep -E "bench\]|rror" $O/bench_n2_weak.err | tail -4
done
echo "== N=2 strong"; BENCH_ALL_RANKS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --scaling strong > $O/bench_n2_strong.json 2> $O/bench_n2_strong.err; python -c "import json; l=json.load(open('$O/bench_n2_strong.json')); print('N=2 strong ms', l['ms_per_step'], 'value', l['value'])"; grep -E "bench\]|rror" $O/bench_n2_strong.err | tail -4
echo "== N=2 weak with e2e"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2 e2e', l['e2e'])"
echo "== N=2 C5 (general partition)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --config C5 --size 32 --steps 5 --no-e2e > $O/bench_n2_C5.json 2> $O/bench_n2_C5.err; tail -c 700 $O/bench_n2_C5.json; tail -3 $O/bench_n2_C5.err
} > $O/session8.log 2>&1
tail -70 $O/session8.log
