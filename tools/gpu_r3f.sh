set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_gpu.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r3f_tests.log
F="--workload gnn_gae --steps 32 --warmup 6 --no-cpu-baseline --no-large --no-e2e"
python bench.py $F > gpurun_out/r3f_gnn.json 2> gpurun_out/r3f_gnn.err
LM_B200_LIB=variants/liblm_gnn_scaled_f0.so python bench.py $F > gpurun_out/r3f_gnn_f0.json 2>/dev/null
LM_B200_LIB=variants/liblm_gnn_unscaled.so python bench.py $F > gpurun_out/r3f_gnn_unscaled.json 2>/dev/null
cat gpurun_out/r3f_tests.log
for f in gpurun_out/r3f_gnn*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['ms_per_step']*1e3,2), {k: round(v*1e3,2) for k,v in d['kernel_ms'].items()})"; done
