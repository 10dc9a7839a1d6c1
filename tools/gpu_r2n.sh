set -x
mkdir -p gpurun_out
python tools/mlp_variants.py locomanipulationrl_b200/liblm_b200.so variants/*.so > gpurun_out/r2n_mlp_variants.log 2>&1
LM_B200_LIB=variants/liblm_diag_both.so python tools/mlp_timeline.py > gpurun_out/r2n_mlp_timeline_both.log 2>&1
cat gpurun_out/r2n_mlp_variants.log gpurun_out/r2n_mlp_timeline_both.log
