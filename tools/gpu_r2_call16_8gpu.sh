#!/bin/bash
# Round 2, GPU call 16 (8 GPUs): correctness of the partitions at 8 ranks, weak + strong scaling of C2, C3 at 128^3 over 8 GPUs
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== dist checks, 8 ranks"
timeout 300 $TR --nproc-per-node 8 --master-port 29601 tests/dist_check.py 6 2>&1 | grep -E "DIST_CHECK|rror" | tail -3
timeout 300 $TR --nproc-per-node 8 --master-port 29602 tests/dist_check_general.py stokes_p2p1_tet 6 2>&1 | grep -E "DIST_CHECK|rror" | tail -3
timeout 300 $TR --nproc-per-node 8 --master-port 29603 tests/dist_check_general.py laplace_q1_hex 16 2>&1 | grep -E "DIST_CHECK|rror" | tail -3
echo "== C2 weak N=8"; BENCH_ALL_RANKS=1 timeout 600 $TR --nproc-per-node 8 --master-port 29604 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e > $O/bench_n8_weak.json 2> $O/bench_n8_weak.err; python -c "import json; l=json.load(open('$O/bench_n8_weak.json')); print('N=8 weak ms', l['ms_per_step'], 'value', l['value'])"; grep -E "bench\]|rror" $O/bench_n8_weak.err | tail -9
echo "== C2 strong N=8"; BENCH_ALL_RANKS=1 timeout 600 $TR --nproc-per-node 8 --master-port 29605 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --scaling strong > $O/bench_n8_strong.json 2> $O/bench_n8_strong.err; python -c "import json; l=json.load(open('$O/bench_n8_strong.json')); print('N=8 strong ms', l['ms_per_step'], 'value', l['value'])"; grep -E "bench\]|rror" $O/bench_n8_strong.err | tail -9
echo "== C2 strong N=4"; timeout 600 $TR --nproc-per-node 4 --master-port 29606 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e --scaling strong > $O/bench_n4_strong.json 2> $O/bench_n4_strong.err; python -c "import json; l=json.load(open('$O/bench_n4_strong.json')); print('N=4 strong ms', l['ms_per_step'], 'value', l['value'])"; grep -E "rror" $O/bench_n4_strong.err | tail -3
echo "== C3 128^3 over 8 GPUs"; timeout 900 $TR --nproc-per-node 8 --master-port 29607 bench.py --gpus 8 --config C3 --size 128 --steps 3 --no-e2e > $O/bench_n8_C3.json 2> $O/bench_n8_C3.err; tail -c 1200 $O/bench_n8_C3.json; grep -E "rror" $O/bench_n8_C3.err | tail -3
} > $O/session16.log 2>&1
tail -60 $O/session16.log
