#!/bin/bash
# A/B: first TMEM chunk of every product fetched ahead of the barrier in front of it (default) vs not (B2T_PCG_TM_PRE=0)
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -x --tb=short -p no:cacheprovider -k "tensor_memory" 2>&1 | tail -3
show() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        it = d['iterations']['pcg_iters_per_instance'] * d['config']['batch_per_gpu']
        print('$1 value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'pcg ns/inst-iter %.2f' % (r['kernel_seconds_per_step']['pcg'] * 1e9 / it), 'frac %.3f' % r['frac'], {k: round(v*1e3,2) for k,v in r['kernel_seconds_per_step'].items()})
"; }
for A in ${PRES:-0 1 0 1}; do B2T_PCG_TM_PRE=$A timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | show "pre=$A"; done
