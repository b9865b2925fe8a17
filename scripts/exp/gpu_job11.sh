python -m pytest tests/test_gpu_parity.py -x -q -k "rescale_backward or training_tail" 2>&1 | tail -3
python scripts/bench_aux.py train 2>&1 | grep -E "rescale x2 bwd|vecint bwd"
python scripts/bench_train_tail.py | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['items_per_s'], d['frac_of_peak_per_gpu'])"
