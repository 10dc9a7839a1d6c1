timeout 900 python -m pytest tests/test_noise.py tests/test_cabi.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t.log
