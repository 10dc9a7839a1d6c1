set -x
mkdir -p gpurun_out
python tools/mlp_variants.py locomanipulationrl_b200/liblm_b200.so variants/liblm_sf1.so variants/liblm_sf2.so > gpurun_out/r3o_mlp_variants.log 2>&1
cat gpurun_out/r3o_mlp_variants.log
