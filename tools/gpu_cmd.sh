set -x
timeout 600 python -m pytest tests/test_policy_gpu.py -m gpu -q -x -k "gnn" 2>&1 | tail -5
timeout 120 python tools/gnn_timeline.py 65536
