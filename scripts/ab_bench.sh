#!/bin/bash
# One short headline-bench line per variant, for A/B runs on one GPU box:
#   scripts/ab_bench.sh label [ENV=value ...]      e.g.  scripts/ab_bench.sh nobar TC_ROT64=4
#   scripts/ab_bench.sh other TC_B200_LIB=$PWD/scratch_libs/libtc_variant.so
# prints: label, chain-steps/s (device resident, end to end), ms per period, Jacobi ms per layer launch, step shares, mean sweeps
label=$1; shift
env "$@" python bench.py --steps 6 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('$label', round(d['value'], 2), 'e2e', round(d['e2e']['value'], 2), 'ms', round(d['ms_per_step'], 2), 'jac_ms', round(r['jacobi_ms_per_launch'], 2), r['step_share'], 'sweeps', d['svd_flags']['mean_sweeps_large'])
"
