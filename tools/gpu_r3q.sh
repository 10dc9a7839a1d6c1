set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout.py tests/test_mlp_train.py tests/test_ppo.py -m gpu -q 2>&1 | tail -2 > gpurun_out/r3q_tests.log
python bench.py --steps 48 --warmup 6 --no-cpu-baseline --no-large --no-e2e --no-kernel-times > gpurun_out/r3q_bench.json 2>/dev/null
cat gpurun_out/r3q_tests.log
python -c "
import json; d=json.load(open('gpurun_out/r3q_bench.json')); print(round(d['ms_per_step']*1e3,2), d['train'])"
