set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_mlp_train.py tests/test_bench_shapes_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2m_tests.log
python tools/mlp_variants.py locomanipulationrl_b200/liblm_b200.so variants/*.so > gpurun_out/r2m_mlp_variants.log 2>&1
python tools/mlp_timeline.py > gpurun_out/r2m_mlp_timeline.log 2>&1
cat gpurun_out/r2m_tests.log gpurun_out/r2m_mlp_variants.log gpurun_out/r2m_mlp_timeline.log
