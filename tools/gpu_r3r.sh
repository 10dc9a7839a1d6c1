set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bench_shapes_gpu.py -m gpu -q -x -k "early_reset_equals or deferred_finish" 2>&1 | tail -8 > gpurun_out/r3r_tests.log
cat gpurun_out/r3r_tests.log
