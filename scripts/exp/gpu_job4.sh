python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; tail -3 gpurun_out/r2c_bench_n1.err; python -c "
import json
d=json.load(open('gpurun_out/r2c_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['e2e'])
print(json.dumps(d['configs'],indent=1))
print({k:(round(v['ms_per_launch'],4), round(v['frac_of_peak'],3)) for k,v in d['kernels'].items()})"
