python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_pipelines_gpu.py -x -q -k "compose or two_steps" 2>&1 | tail -4
python scripts/bench_aux.py misc 2>&1 | grep -E "compose"
python scripts/bench_two_step.py | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_pass'], d['subjects_per_s'], d['frac_of_peak_per_gpu'])"
