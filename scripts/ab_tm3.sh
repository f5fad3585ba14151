#!/bin/bash
# A/B: k_pcg_tm specialised for N = 64 / Euler (default) vs the run-time instantiation (B2T_PCG_TM_GENERIC=1)
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -x --tb=short -p no:cacheprovider -k "tensor_memory" 2>&1 | tail -3
show() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        it = d['iterations']['pcg_iters_per_instance'] * d['config']['batch_per_gpu']
        print('$1 value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'pcg ns/inst-iter %.2f' % (r['kernel_seconds_per_step']['pcg'] * 1e9 / it), 'frac %.3f' % r['frac'], {k: round(v*1e3,2) for k,v in r['kernel_seconds_per_step'].items()})
"; }
for G in ${GEN:-1 0 1 0}; do B2T_PCG_TM_GENERIC=$G timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | show "generic=$G"; done
