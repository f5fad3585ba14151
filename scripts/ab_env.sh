#!/bin/bash
# A/B of any environment switch of the library on the default bench workload, alternating runs:
#   gpurun --timeout 900 -- 'bash scripts/ab_env.sh B2T_PCG_TM_NBR "1 0 1 0"'        (also B2T_PCG_TM_PRE, B2T_PCG_TM_GENERIC, B2T_PCG_VARIANT, B2T_LS_PAR ...)
# EXTRA="--batch 2048 --limits 0" adds bench.py arguments; TESTS=<pytest -k expression> runs those GPU tests first with the first value set
VAR=${1:?variable}; VALS=${2:?values}
show() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        it = d['iterations']['pcg_iters_per_instance'] * d['config']['batch_per_gpu']
        print('$1 value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'pcg ns/inst-iter %.2f' % (r['kernel_seconds_per_step']['pcg'] * 1e9 / it), 'frac %.3f' % r['frac'], {k: round(v*1e3,2) for k,v in r['kernel_seconds_per_step'].items()})
"; }
if [ -n "$TESTS" ]; then env $VAR=${VALS%% *} timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -x --tb=short -p no:cacheprovider -k "$TESTS" 2>&1 | tail -3; fi
for A in $VALS; do env $VAR=$A timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline $EXTRA 2>/dev/null | show "$VAR=$A"; done
