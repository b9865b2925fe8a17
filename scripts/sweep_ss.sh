#!/bin/bash
# tuning sweep of the plane-marching SS kernel (run under gpurun); one bench.py line per configuration
out=gpurun_out/sweep_ss.jsonl
: > $out
run() {  # label, env...
  label=$1; shift
  echo "== $label" >> $out
  env "$@" python bench.py --steps 8 --warmup 3 --no-e2e 2>>gpurun_out/sweep_ss.err | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    k = d['kernels']
    print(json.dumps({'ms_per_step': d['ms_per_step'], 'value': d['value'], **{n.split('(')[0]: [round(v['ms_per_launch'], 4), round(v['frac_of_peak'], 3)] for n, v in k.items()}}))
" >> $out
}
for a in "$@"; do run "$a" $a; done
cat $out
