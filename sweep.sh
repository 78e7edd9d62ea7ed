timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rows in 400 300; do echo "== ROWS=$rows"; ISL_PATCH_ROWS=$rows timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e 2>&1 | grep -E "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms_step %.2f kernel_ms %.2f frac %.3f perturbed %.2f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['ms_per_step_perturbed_mesh']))
"; done
