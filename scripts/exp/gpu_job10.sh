python -m pytest tests -x -q -m gpu -k "jacobian or jacdet" 2>&1 | tail -3
python scripts/bench_jacobian.py
python scripts/bench_aux.py jac
