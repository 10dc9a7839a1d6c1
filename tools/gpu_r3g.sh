set -x
mkdir -p gpurun_out
python tools/sanitize_run.py > gpurun_out/r3g_plain.log 2>&1 &&
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_run.py > gpurun_out/r3g_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r3g_memcheck.log
cat gpurun_out/r3g_plain.log; tail -15 gpurun_out/r3g_memcheck.log
