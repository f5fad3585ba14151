#!/bin/bash
# A/B of the PCG kernels on the limits / no-limits workloads (2048 instances): k_pcg3 (B2T_PCG_VARIANT=3) vs k_pcg4 (default, 5)
one() { timeout 600 python bench.py --batch ${BATCH:-2048} --steps 2 --limits $1 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        it = d['iterations']['pcg_iters_per_instance'] * d['config']['batch_per_gpu']
        print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'pcg ns/inst-iter %.2f' % (r['kernel_seconds_per_step']['pcg'] * 1e9 / it), 'frac %.3f' % r['frac'], {k: round(v*1e3,2) for k,v in r['kernel_seconds_per_step'].items()})
    elif 'rror' in l: print(l.strip()[:300])
"; }
for V in ${VARIANTS:-3 5}; do for L in ${LIMITS:-0 1}; do echo "--- variant $V limits $L"; B2T_PCG_VARIANT=$V one $L; done; done
