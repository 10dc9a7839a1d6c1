timeout 900 python -m pytest tests/test_rollout.py tests/test_ppo.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/t.log
python tools/update_bench.py > gpurun_out/update_bench.log 2>&1
