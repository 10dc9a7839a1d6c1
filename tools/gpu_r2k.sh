set -x
mkdir -p gpurun_out
python tools/mlp_variants.py locomanipulationrl_b200/liblm_b200.so variants/*.so > gpurun_out/r2k_mlp_variants.log 2>&1
LM_B200_LIB=variants/liblm_ef_relu.so python tools/mlp_timeline.py > gpurun_out/r2k_mlp_timeline_ef_relu.log 2>&1
cat gpurun_out/r2k_mlp_variants.log gpurun_out/r2k_mlp_timeline_ef_relu.log
