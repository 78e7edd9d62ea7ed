run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e 2>&1 | grep -E "^{" | tee gpurun_out/bench_g.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms_step %.2f kernel_ms %.2f frac %.3f perturbed %.2f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['ms_per_step_perturbed_mesh']))
"; }
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run ISL_DBG=0
