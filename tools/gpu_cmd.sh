timeout 900 python -m pytest tests/test_rollout.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/t.log
timeout 300 python bench.py --workload gnn_gae --steps 64 --warmup 6 --no-cpu-baseline --no-e2e --no-large > gpurun_out/gnn8.json 2> gpurun_out/gnn8.err
