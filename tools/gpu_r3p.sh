set -x
mkdir -p gpurun_out
F="--steps 60 --warmup 6 --no-cpu-baseline --no-train --no-large --no-kernel-times"
for sp in 0.5 0.625 0.75 0.875; do LM_BENCH_E2E_SPLIT=$sp python bench.py $F > gpurun_out/r3p_split_$sp.json 2>/dev/null; done
for f in gpurun_out/r3p_split_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', 'e2e', round(d['e2e']['ms_per_step'],4))"; done
