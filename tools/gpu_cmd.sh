timeout 900 python -m pytest tests/test_noise.py tests/test_ppo.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/t.log
python tools/update_bench.py > gpurun_out/update_bench.log 2>&1
B="python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-graphs --no-e2e --no-large --no-kernel-times"
$B > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v5.csv $B > gpurun_out/ncu1.log 2>&1
G="python bench.py --workload gnn_gae --steps 32 --warmup 3 --no-cpu-baseline --no-graphs --no-e2e --no-large --no-kernel-times"
$G > gpurun_out/plain_gnn2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v5_gnn.csv $G > gpurun_out/ncu3.log 2>&1
