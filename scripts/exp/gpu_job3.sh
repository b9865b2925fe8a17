nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,temperature.gpu --format=csv,noheader
for i in 1 2; do python bench.py --steps 10 --warmup 3 --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:(round(v['ms_per_launch'],4), round(v['frac_of_peak'],3)) for k,v in d['kernels'].items()}, d['clocks'])"; done
