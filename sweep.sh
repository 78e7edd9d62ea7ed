timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for ws in 0 1; do echo "== WS=$ws"; ISL_PROF=1 ISL_PATCH_WS=$ws timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e 2>&1 | grep -E "isl-prof|^{" | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('[isl'): print(l.strip()[:250])
    if l.startswith('{'):
        d=json.loads(l); print('ms_step %.2f kernel_ms %.2f frac %.3f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))
"; done
