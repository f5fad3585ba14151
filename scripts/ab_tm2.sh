#!/bin/bash
# second A/B round: long horizon (N = 128) with k_pcg_tm<512> vs k_pcg3<512>, and the all-trials-at-once line search in the bulk
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -q -x --tb=short -p no:cacheprovider -k "tensor_memory" 2>&1 | tail -5
show() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        it = d['iterations']['pcg_iters_per_instance'] * d['config']['batch_per_gpu']
        print('$1 value', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'pcg ns/inst-iter %.2f' % (r['kernel_seconds_per_step']['pcg'] * 1e9 / it), 'frac %.3f' % r['frac'], {k: round(v*1e3,2) for k,v in r['kernel_seconds_per_step'].items()}, d.get('passes', {}).get('passes_per_step'))
"; }
for V in 3 8; do B2T_BENCH_KNOTS=128 B2T_PCG_VARIANT=$V timeout 600 python bench.py --batch 4096 --steps 2 --warmup 1 --no-cpu-baseline 2>$O/ab_tm2_N128_$V.err | tee $O/ab_tm2_N128_$V.json | show "N128 variant $V"; tail -2 $O/ab_tm2_N128_$V.err; done
B2T_LS_PAR=1000000 timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | show "LS_PAR all"
timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | show "default"
