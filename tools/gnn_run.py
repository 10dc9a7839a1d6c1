"""Runs the tensor-core GNN forward a few times at 65,536 envs (ncu target).  python tools/gnn_run.py [n_envs] [launches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from locomanipulationrl_b200.policy.gnn import GnnActorCritic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
m = GnnActorCritic(device="cuda:0", precision="bf16")
obs = torch.randn(n, 64, device="cuda")
for _ in range(k):
    mu, value, _ = m(obs)
torch.cuda.synchronize()
print("ok", float(mu.abs().max()), float(value.abs().max()))
