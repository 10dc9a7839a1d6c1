# What a round-end check on a B200 box runs (gpurun -- 'mkdir -p gpurun_out; bash tools/gpu_cmd.sh'): GPU parity suite, smoke, default bench.
set -x
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/t.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/final_default.json 2> gpurun_out/final_default.err
python bench.py --workload gnn_gae --steps 192 --warmup 12 --no-cpu-baseline > gpurun_out/final_gnn.json 2> gpurun_out/final_gnn.err
