timeout 600 python -m pytest tests/test_mlp_train.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/t.log
