set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2r_tests.log
F="--steps 96 --warmup 6 --no-cpu-baseline --no-train --no-large --no-e2e"
python bench.py $F > gpurun_out/r2r_bench_defer.json 2> gpurun_out/r2r_bench_defer.err
LM_BENCH_INLINE_FINISH=1 python bench.py $F > gpurun_out/r2r_bench_inline.json 2> gpurun_out/r2r_bench_inline.err
tail -12 gpurun_out/r2r_tests.log; tail -3 gpurun_out/r2r_bench_defer.err
for f in gpurun_out/r2r_bench_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['ms_per_step']*1e3,2), d['kernel_ms'], d['roofline']['kernel'], round(d['roofline']['frac'],3))"; done
