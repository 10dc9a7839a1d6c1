timeout 1200 python -m pytest tests/test_task_gpu.py tests/test_cabi.py -m gpu -q -x -k "CustomController or custom_controller or cabi" 2>&1 | tail -25 > gpurun_out/t.log
