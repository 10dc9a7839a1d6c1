timeout 900 python -m pytest tests/test_ppo.py tests/test_rollout.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t.log
