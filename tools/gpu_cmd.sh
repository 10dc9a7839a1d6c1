timeout 1200 python -m pytest tests/test_task_gpu.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/t.log
python bench.py --workload manip16k --steps 192 --warmup 12 --no-cpu-baseline --no-e2e --no-large > gpurun_out/final_manip.json 2> gpurun_out/final_manip.err
