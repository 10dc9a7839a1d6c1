set -x
mkdir -p gpurun_out
F="--steps 96 --warmup 6 --no-cpu-baseline --no-train --no-large --no-e2e"
python bench.py $F > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
LM_B200_LIB=variants/liblm_noreset.so python bench.py $F > gpurun_out/r2s_bench_noreset.json 2> gpurun_out/r2s_bench_noreset.err
python bench.py $F --workload gnn_gae > gpurun_out/r2s_gnn.json 2> gpurun_out/r2s_gnn.err
python bench.py $F --workload manip16k > gpurun_out/r2s_manip.json 2> gpurun_out/r2s_manip.err
python tools/gnn_timeline.py > gpurun_out/r2s_gnn_timeline.log 2>&1
for f in gpurun_out/r2s_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['ms_per_step']*1e3,2), {k: round(v*1e3,2) for k,v in d['kernel_ms'].items()}, d['roofline']['kernel'], round(d['roofline']['frac'],3))"; done
tail -20 gpurun_out/r2s_gnn_timeline.log
