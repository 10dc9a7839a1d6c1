# round 2: ncu evidence for the final kernels (launch list + full capture), one GPU
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-graphs --no-e2e --no-train --no-large --no-kernel-times"
$CMD > gpurun_out/r2y_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2y_launches.csv $CMD > gpurun_out/r2y_ncu1.log 2>&1
$CMD > gpurun_out/r2y_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_mlp_forward|k_post_lean|k_pre' -s 30 -c 6 -o gpurun_out/r2y_prof $CMD > gpurun_out/r2y_ncu2.log 2>&1
CMD2="python bench.py --workload gnn_gae --steps 16 --warmup 3 --no-cpu-baseline --no-graphs --no-e2e --no-large --no-kernel-times"
$CMD2 > gpurun_out/r2y_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_gnn_forward_tc|k_rollout_write' -s 8 -c 4 -o gpurun_out/r2y_prof_gnn $CMD2 > gpurun_out/r2y_ncu3.log 2>&1
ls -la gpurun_out/r2y_*; tail -3 gpurun_out/r2y_ncu2.log
