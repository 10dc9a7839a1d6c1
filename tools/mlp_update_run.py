"""Runs the tcgen05 PPO gradient chain of the MLP a few times at a 262,144-sample minibatch (ncu target).  python tools/mlp_update_run.py [B] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from locomanipulationrl_b200.policy.mlp import MlpActorCritic  # noqa: E402
from locomanipulationrl_b200.ppo import MlpPpoGradients  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda:0"
torch.manual_seed(0)
model = MlpActorCritic(device=dev)
g = torch.Generator(device=dev).manual_seed(1)
obs = torch.randn(B, 64, device=dev, generator=g).clamp(-5, 5)
act, old_mu = torch.randn(B, 12, device=dev, generator=g) * 0.3, torch.randn(B, 12, device=dev, generator=g) * 0.3
nlp, adv, ov, ret = (torch.randn(B, device=dev, generator=g) for _ in range(4))
ls = torch.zeros(12, device=dev)
grads = MlpPpoGradients(model, max_batch=B)
for _ in range(reps):
    stats = grads(obs, act, nlp + 11.0, adv, ov, ret, old_mu, ls)
torch.cuda.synchronize()
print("ok", [round(float(x), 4) for x in stats[:6]])
