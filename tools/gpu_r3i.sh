# round 2: repeatability of the GPU suite (the new kernels run concurrently with others: look for flakiness), then the final ncu evidence
set -x
mkdir -p gpurun_out
for i in 1 2 3 4 5; do timeout 600 python -m pytest tests/test_bench_shapes_gpu.py tests/test_task_gpu.py -m gpu -q -x 2>&1 | tail -1 >> gpurun_out/r3i_repeat.log; done
for i in 1 2; do timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -1 >> gpurun_out/r3i_repeat.log; done
python __graft_entry__.py smoke >> gpurun_out/r3i_repeat.log 2>&1
cat gpurun_out/r3i_repeat.log
CMD="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-graphs --no-e2e --no-train --no-large --no-kernel-times"
$CMD > gpurun_out/r3i_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r3i_launches.csv $CMD > gpurun_out/r3i_ncu1.log 2>&1
$CMD > gpurun_out/r3i_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_mlp_forward|k_post_lean|k_pre' -s 30 -c 6 -o gpurun_out/r3i_prof $CMD > gpurun_out/r3i_ncu2.log 2>&1
tail -2 gpurun_out/r3i_ncu2.log
