"""2-GPU tier (SURVEY §4 last row): runs tests/multi_gpu_worker.py under torchrun on two GPUs of the box.  Skipped with fewer than
two GPUs (the default single-GPU test box); run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_collectives_and_ppo_update():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(res.stdout[-12000:])
    sys.stderr.write(res.stderr[-3000:])
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):                          # the full transcript, for the record
        with open(os.path.join(out_dir, "multi_gpu_worker.log"), "w") as f:
            f.write(res.stdout + "\n---- stderr ----\n" + res.stderr[-6000:])
    assert res.returncode == 0, "a rank reported failures (see the captured output)"
    assert "FAIL" not in res.stdout
