timeout 600 python -m pytest tests/test_mlp_train.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/t.log
python tools/update_bench.py 2>&1 | tail -3 > gpurun_out/update_bench2.log
