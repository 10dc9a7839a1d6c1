set -x
mkdir -p gpurun_out
F="--steps 96 --warmup 6 --no-cpu-baseline --no-train --no-large --no-e2e"
python bench.py $F > gpurun_out/r3h_base.json 2>/dev/null
for v in 800 1500 2500; do LM_B200_LIB=variants/liblm_stagger$v.so python bench.py $F > gpurun_out/r3h_stagger$v.json 2>/dev/null; done
for f in gpurun_out/r3h_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['ms_per_step']*1e3,2), {k: round(v*1e3,2) for k,v in d['kernel_ms'].items()})"; done
