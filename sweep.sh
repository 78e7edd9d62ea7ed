timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for fast in 1 3; do echo "== FAST=$fast"; ISL_Q1_FAST=$fast timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e 2>&1 | grep -E "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms_step %.2f kernel_ms %.2f frac %.3f perturbed %.2f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['ms_per_step_perturbed_mesh']))
"; done
