set -x
python bench.py --steps 2 --warmup 3 --batch 32 --no-e2e > gpurun_out/r2_b32_plain.json 2> gpurun_out/r2_b32_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_b32.csv python bench.py --steps 2 --warmup 3 --batch 32 --no-e2e > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_ss_march|k_warp_brick_var|k_upsample3_march' -s 36 -c 12 -o gpurun_out/prof_r2_b32 -f python bench.py --steps 2 --warmup 3 --batch 32 --no-e2e > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/prof_r2_b32.ncu-rep
