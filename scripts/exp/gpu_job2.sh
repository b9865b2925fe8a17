python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "fused or headline or vxmdense or graphed" 2>&1 | tail -5
for m in 1 5; do echo "== MINB=$m"; DFM_WBV_MINB=$m python bench.py --steps 10 --warmup 3 --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:(round(v['ms_per_launch'],4), round(v['frac_of_peak'],3)) for k,v in d['kernels'].items()})"; done
