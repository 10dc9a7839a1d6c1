# round 2: full validation of the tree (GPU suite, smoke, the three bench workloads with all blocks)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r3n_tests.log
python __graft_entry__.py smoke > gpurun_out/r3n_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r3n_bench.json 2> gpurun_out/r3n_bench.err
timeout 600 python bench.py --workload manip16k --no-cpu-baseline > gpurun_out/r3n_manip.json 2> gpurun_out/r3n_manip.err
timeout 600 python bench.py --workload gnn_gae --no-cpu-baseline > gpurun_out/r3n_gnn.json 2> gpurun_out/r3n_gnn.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r3n_ref.json 2> gpurun_out/r3n_ref.err
tail -5 gpurun_out/r3n_tests.log; tail -2 gpurun_out/r3n_smoke.log; tail -3 gpurun_out/r3n_bench.err
for f in gpurun_out/r3n_bench.json gpurun_out/r3n_manip.json gpurun_out/r3n_gnn.json; do python -c "
import json; d=json.load(open('$f')); print('$f', '%.4g'%d['value'], round(d['ms_per_step']*1e3,2), {k: round(v*1e3,2) for k,v in d['kernel_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],3), d['roofline']['kernel'], round(d['roofline']['frac'],3), 'train', d.get('train') and round(d['train']['ms_per_iteration'],2))"; done
cat gpurun_out/r3n_ref.json | head -c 600
