python -m pytest tests/test_pipelines_gpu.py -x -q 2>&1 | tail -5
python scripts/probe_e2e.py 2>&1 | tail -9
