timeout 600 python -m pytest tests/test_rollout.py tests/test_cabi.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/t.log
