timeout 600 python -m pytest tests/test_policy_gpu.py -m gpu -q -x -k gnn 2>&1 | tail -3 > gpurun_out/t.log
timeout 120 python tools/gnn_timeline.py 65536 > gpurun_out/tl_base.log 2>&1
timeout 300 python bench.py --workload gnn_gae --steps 64 --warmup 6 --no-cpu-baseline --no-e2e --no-large > gpurun_out/gnn7.json 2> gpurun_out/gnn7.err
