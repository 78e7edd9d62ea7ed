run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --no-cpu-baseline --no-e2e 2>&1 | grep -E "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['ms_per_step'], d['clocks']['power_w_max'], d['roofline']['kernel_ms'])
"; }
echo "== no exchange"; BENCH_NO_EXCHANGE=1 run 29601
echo "== NCCL channels limited"; NCCL_MAX_NCHANNELS=2 NCCL_MAX_P2P_NCHANNELS=2 NCCL_MIN_P2P_NCHANNELS=1 run 29602
echo "== default"; run 29603
