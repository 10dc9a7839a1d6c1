set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_bench_shapes_gpu.py tests/test_policy_gpu.py tests/test_mlp_train.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r3b_tests.log
F="--steps 96 --warmup 6 --no-cpu-baseline --no-train --no-large --no-e2e"
timeout 300 python bench.py $F > gpurun_out/r3b_bench_default.json 2> gpurun_out/r3b_bench_default.err
python tools/mlp_variants.py locomanipulationrl_b200/liblm_b200.so > gpurun_out/r3b_mlp_variants.log 2>&1
cat gpurun_out/r3b_tests.log gpurun_out/r3b_mlp_variants.log
for f in gpurun_out/r3b_bench_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['ms_per_step']*1e3,2), {k: round(v*1e3,2) for k,v in d['kernel_ms'].items()})"; done
