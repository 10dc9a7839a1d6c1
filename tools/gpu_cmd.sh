python tools/update_bench.py > gpurun_out/update_bench.log 2>&1
