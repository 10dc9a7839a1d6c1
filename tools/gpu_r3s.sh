set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bench_shapes_gpu.py -m gpu -q -x -k "multi_step_graph" 2>&1 | tail -3 > gpurun_out/r3s_tests.log
F="--steps 96 --warmup 6 --no-cpu-baseline --no-train --no-large --no-e2e"
python bench.py $F > gpurun_out/r3s_bench48.json 2>/dev/null
LM_BENCH_GRAPH_STEPS=6 python bench.py $F > gpurun_out/r3s_bench6.json 2>/dev/null
python bench.py --workload gnn_gae --steps 32 --warmup 6 --no-cpu-baseline --no-large --no-e2e --no-kernel-times > gpurun_out/r3s_gnn.json 2>/dev/null
cat gpurun_out/r3s_tests.log
for f in gpurun_out/r3s_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['ms_per_step']*1e3,2), d['warmup'], d['steps'])"; done
