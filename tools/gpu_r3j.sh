set -x
mkdir -p gpurun_out
python tools/mlp_variants.py locomanipulationrl_b200/liblm_b200.so variants/liblm_pin.so > gpurun_out/r3j_mlp_variants.log 2>&1
F="--steps 96 --warmup 6 --no-cpu-baseline --no-train --no-large --no-e2e"
LM_B200_LIB=variants/liblm_pin.so python bench.py $F > gpurun_out/r3j_pin.json 2>/dev/null
cat gpurun_out/r3j_mlp_variants.log
python -c "
import json; d=json.load(open('gpurun_out/r3j_pin.json')); print(round(d['ms_per_step']*1e3,2), {k: round(v*1e3,2) for k,v in d['kernel_ms'].items()})"
