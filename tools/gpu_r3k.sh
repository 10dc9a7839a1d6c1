set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_task_gpu.py tests/test_bench_shapes_gpu.py -m gpu -q -x 2>&1 | tail -2 > gpurun_out/r3k_tests.log
python tools/post_timeline.py 65536 > gpurun_out/r3k_post_timeline.log 2>&1
F="--steps 96 --warmup 6 --no-cpu-baseline --no-train --no-large --no-e2e"
python bench.py $F > gpurun_out/r3k_bench.json 2>/dev/null
cat gpurun_out/r3k_tests.log; head -3 gpurun_out/r3k_post_timeline.log
python -c "
import json; d=json.load(open('gpurun_out/r3k_bench.json')); print(round(d['ms_per_step']*1e3,2), {k: round(v*1e3,2) for k,v in d['kernel_ms'].items()})"
