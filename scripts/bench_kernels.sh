#!/bin/bash
# usage: scripts/bench_kernels.sh [env assignments...]  -> per-kernel ms / GB/s / fraction
env "$@" python bench.py --steps 10 --warmup 3 --no-e2e 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('  %.3e vox/s  %.3f ms/step' % (d['value'], d['ms_per_step']))
for k,v in d['kernels'].items(): print('  %-32s %.4f ms  %5.0f GB/s  %.3f' % (k, v['ms_per_launch'], v['achieved_gbs'], v['frac_of_peak']))
"
