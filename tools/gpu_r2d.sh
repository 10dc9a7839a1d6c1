set -x
mkdir -p gpurun_out
F="--steps 96 --warmup 6 --no-e2e --no-cpu-baseline --no-train --no-large"
python bench.py $F > gpurun_out/r2d_default.json 2> gpurun_out/r2d_default.err
LM_BENCH_INLINE_FINISH=1 python bench.py $F > gpurun_out/r2d_inlinefinish.json 2>/dev/null
LM_NO_PDL=1 python bench.py $F > gpurun_out/r2d_nopdl.json 2>/dev/null
LM_NO_PDL=1 LM_BENCH_NO_OVERLAP=1 python bench.py $F > gpurun_out/r2d_nopdl_nooverlap.json 2>/dev/null
LM_BENCH_EARLY_RESET=1 python bench.py $F > gpurun_out/r2d_early.json 2>/dev/null
LM_BENCH_EARLY_RESET=1 LM_BENCH_INLINE_FINISH=1 python bench.py $F > gpurun_out/r2d_early_inline.json 2>/dev/null
LM_BENCH_NO_OVERLAP=1 python bench.py $F --no-graphs > gpurun_out/r2d_eager_nooverlap.json 2>/dev/null
python bench.py $F --no-graphs > gpurun_out/r2d_eager_overlap.json 2>/dev/null
timeout 600 python -m pytest tests/test_bench_shapes_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2d_tests.log
for f in gpurun_out/r2d_*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', round(d['ms_per_step']*1e3,2), d.get('kernel_ms'))"; done
cat gpurun_out/r2d_tests.log
