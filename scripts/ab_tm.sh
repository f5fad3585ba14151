#!/bin/bash
# A/B of k_pcg_tm (B2T_PCG_VARIANT=8: matrices in tensor memory, two instances per SM) against k_pcg3 (3) on the default workload
#   gpurun --timeout 900 -- 'bash scripts/ab_tm.sh'
O=gpurun_out; mkdir -p $O
BATCH=${BATCH:-8192}
if [ -z "$SKIP_TESTS" ]; then timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -x --tb=short -p no:cacheprovider -k "tensor_memory or pcg_kernel" 2>&1 | tail -15; fi
for V in ${VARIANTS:-3 8}; do for L in ${LIMITS:-1 0}; do
B2T_PCG_VARIANT=$V timeout 600 python bench.py --batch $BATCH --steps 3 --warmup 1 --limits $L --no-cpu-baseline 2>$O/ab_tm_${V}_${L}.err | tee $O/ab_tm_${V}_${L}.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        it = d['iterations']['pcg_iters_per_instance'] * d['config']['batch_per_gpu']
        print('variant $V limits $L value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'pcg ns/inst-iter %.2f' % (r['kernel_seconds_per_step']['pcg'] * 1e9 / it), 'frac %.3f' % r['frac'], 'iters/inst %.1f' % d['iterations']['pcg_iters_per_instance'], {k: round(v*1e3,2) for k,v in r['kernel_seconds_per_step'].items()}, d.get('parity_check'))
"; tail -3 $O/ab_tm_${V}_${L}.err; done; done
