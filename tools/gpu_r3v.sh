set -x
mkdir -p gpurun_out
F="--steps 60 --warmup 6 --no-cpu-baseline --no-train --no-large --no-kernel-times"
python bench.py $F > gpurun_out/r3v_base.json 2>/dev/null
LM_BENCH_WC=1 python bench.py $F > gpurun_out/r3v_wc.json 2> gpurun_out/r3v_wc.err
LM_BENCH_WC=1 python bench.py $F > gpurun_out/r3v_wc2.json 2>/dev/null
python bench.py $F > gpurun_out/r3v_base2.json 2>/dev/null
tail -3 gpurun_out/r3v_wc.err
for f in gpurun_out/r3v_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', 'e2e', round(d['e2e']['ms_per_step'],4), 'readback', round(d['e2e']['with_obs_readback']['ms_per_step'],4))"; done
