#!/bin/bash
# A/B on the limits workload (2048 instances): default, k_schur_rows at 3 CTAs / SM
one() { timeout 600 python bench.py --batch 2048 --steps 2 --limits 1 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), {k: round(v*1e3,2) for k,v in r['kernel_seconds_per_step'].items()})
"; }
echo "--- default"; one
echo "--- B2T_SCHUR_MINB=3"; B2T_SCHUR_MINB=3 one
