#!/bin/bash
# Round 2, GPU call 7 (2 GPUs): NCCL exchange inside the engine: correctness (slab + general partition), weak and strong scaling at N=2
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
nvidia-smi --query-gpu=index,name --format=csv
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_multigpu.py -q 2>&1 | tail -15
echo "== N=1 reference point"; timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 ms', l['ms_per_step'], 'value', l['value'])"
echo "== N=2 weak"; ISL_VERBOSE=1 BENCH_ALL_RANKS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > $O/bench_n2_weak.json 2> $O/bench_n2_weak.err; tail -c 1500 $O/bench_n2_weak.json; grep -E "bench\]|exchange overlap|rror" $O/bench_n2_weak.err | tail -8
echo "== N=2 weak, no overlap (torch exchange)"; ISL_TORCH_EXCHANGE=1 BENCH_ALL_RANKS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e 2> $O/bench_n2_torch.err | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2 torch exchange ms', l['ms_per_step'], 'value', l['value'])"; grep -E "bench\]|rror" $O/bench_n2_torch.err | tail -4
echo "== N=2 strong"; BENCH_ALL_RANKS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --scaling strong > $O/bench_n2_strong.json 2> $O/bench_n2_strong.err; tail -c 800 $O/bench_n2_strong.json; grep -E "bench\]|rror" $O/bench_n2_strong.err | tail -4
echo "== N=2 C5 (general partition)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --config C5 --n 32 --steps 5 --no-e2e > $O/bench_n2_C5.json 2> $O/bench_n2_C5.err; tail -c 600 $O/bench_n2_C5.json; tail -3 $O/bench_n2_C5.err
} > $O/session7.log 2>&1
tail -70 $O/session7.log
