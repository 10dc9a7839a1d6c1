# round 2, GPU call G: pipelined host step (one copy stream, one C-ABI call per shard), MLP epilogue variants + timeline
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bench_shapes_gpu.py tests/test_cabi.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2g_tests.log
F="--steps 96 --warmup 6 --no-cpu-baseline --no-train --no-large --no-kernel-times"
for sh in 1 2 4 8; do
  timeout 600 python bench.py $F --e2e-shards $sh > gpurun_out/r2g_e2e_s$sh.json 2> gpurun_out/r2g_e2e_s$sh.err
done
python tools/mlp_variants.py locomanipulationrl_b200/liblm_b200.so variants/liblm_relu.so variants/liblm_f1.so variants/liblm_f2.so variants/liblm_relu_f1.so > gpurun_out/r2g_mlp_variants.log 2>&1
python tools/mlp_timeline.py > gpurun_out/r2g_mlp_timeline.log 2>&1
LM_B200_LIB=variants/liblm_relu.so python tools/mlp_timeline.py > gpurun_out/r2g_mlp_timeline_relu.log 2>&1
cat gpurun_out/r2g_tests.log gpurun_out/r2g_mlp_variants.log
for sh in 1 2 4 8; do python -c "
import json; d=json.load(open('gpurun_out/r2g_e2e_s$sh.json')); e=d['e2e']; print('shards $sh', round(e['ms_per_step'],4), 'ms/step', 'readback', round(e['with_obs_readback']['ms_per_step'],4))"; done
head -12 gpurun_out/r2g_mlp_timeline.log
