set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_mlp_train.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2j_tests.log
python tools/mlp_variants.py locomanipulationrl_b200/liblm_b200.so variants/*.so > gpurun_out/r2j_mlp_variants.log 2>&1
python tools/mlp_timeline.py > gpurun_out/r2j_mlp_timeline.log 2>&1
LM_B200_LIB=variants/liblm_aux0_ring4.so python tools/mlp_timeline.py > gpurun_out/r2j_mlp_timeline_aux0_ring4.log 2>&1
cat gpurun_out/r2j_tests.log gpurun_out/r2j_mlp_variants.log gpurun_out/r2j_mlp_timeline.log gpurun_out/r2j_mlp_timeline_aux0_ring4.log
