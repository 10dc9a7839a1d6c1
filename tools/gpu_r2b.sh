# round 2, GPU call B: full GPU suite (no -x), smoke, default bench
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -80 > gpurun_out/r2b_tests.log
python __graft_entry__.py smoke > gpurun_out/r2b_smoke.log 2>&1
timeout 900 python bench.py --steps 96 --warmup 6 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
LM_BENCH_NO_SINK=1 timeout 900 python bench.py --steps 96 --warmup 6 --no-e2e --no-cpu-baseline --no-train --no-large > gpurun_out/r2b_bench_nosink.json 2> gpurun_out/r2b_bench_nosink.err
tail -30 gpurun_out/r2b_tests.log; tail -3 gpurun_out/r2b_bench.err
