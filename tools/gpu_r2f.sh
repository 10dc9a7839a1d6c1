# round 2, GPU call F (first call of the re-entered session): full GPU suite, smoke, default bench, ncu launch list + full capture
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2f_tests.log
python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 600 python bench.py --workload manip16k --no-cpu-baseline --no-train > gpurun_out/r2f_manip.json 2> gpurun_out/r2f_manip.err
timeout 600 python bench.py --workload gnn_gae --no-cpu-baseline > gpurun_out/r2f_gnn.json 2> gpurun_out/r2f_gnn.err
CMD="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-graphs --no-e2e --no-train --no-large --no-kernel-times"
$CMD > gpurun_out/r2f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu1.log 2>&1
$CMD > gpurun_out/r2f_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_mlp_forward|k_post_lean|k_pre' -s 30 -c 6 -o gpurun_out/r2f_prof $CMD > gpurun_out/r2f_ncu2.log 2>&1
tail -30 gpurun_out/r2f_tests.log; tail -3 gpurun_out/r2f_smoke.log; tail -3 gpurun_out/r2f_bench.err; cat gpurun_out/r2f_bench.json
