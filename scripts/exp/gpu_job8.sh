python -m pytest tests/test_gpu_parity.py -x -q -k "vecint_backward or ss_step_bwd or training_tail or compose_backward" 2>&1 | grep -E "passed|failed|Mismatch|Max abs|Max rel|FAILED" | head
python scripts/exp/bwd_probe.py
