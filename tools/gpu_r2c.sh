set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2c_tests.log
timeout 900 python bench.py --steps 96 --warmup 6 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
LM_BENCH_NO_OVERLAP=1 timeout 900 python bench.py --steps 96 --warmup 6 --no-e2e --no-cpu-baseline --no-train --no-large > gpurun_out/r2c_bench_nooverlap.json 2> gpurun_out/r2c_bench_nooverlap.err
timeout 600 python bench.py --workload gnn_gae --steps 96 --warmup 6 --no-cpu-baseline --no-e2e > gpurun_out/r2c_gnn.json 2> gpurun_out/r2c_gnn.err
tail -30 gpurun_out/r2c_tests.log; tail -3 gpurun_out/r2c_bench.err
