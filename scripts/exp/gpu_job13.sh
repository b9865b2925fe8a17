export DFM_SAN=1
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_metrics_gpu.py tests/test_synth_gpu.py -x -q -k "channelwise or ss_step_bwd_gather or vecint_backward_gather or rescale_backward_separable or metrics_against or histogram_of_constant or overlap_metrics or synth_intensity or conv1d or norm_gamma" 2>&1 | tail -15
echo "sanitizer rc=$?"
