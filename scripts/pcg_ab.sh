#!/bin/bash
# A/B of the PCG kernel on the no-limits and limits workloads (2048 instances); prints solves/s and per-kernel ms
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_solve.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3; fi
for L in 0 1; do timeout 600 python bench.py --batch 2048 --steps 2 --limits $L --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        it = d['iterations']['pcg_iters_per_instance'] * d['config']['batch_per_gpu']
        print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'pcg ns/inst-iter %.2f' % (r['kernel_seconds_per_step']['pcg'] * 1e9 / it), 'iters/inst %.1f' % d['iterations']['pcg_iters_per_instance'], {k: round(v*1e3,2) for k,v in r['kernel_seconds_per_step'].items()})
"; done
