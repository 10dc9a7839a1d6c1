set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-large --no-kernel-times"
$CMD > gpurun_out/r3l_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r3l_launches_train.csv $CMD > gpurun_out/r3l_ncu1.log 2>&1
tail -2 gpurun_out/r3l_ncu1.log | head -c 300
