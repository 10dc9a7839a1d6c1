"""Diagnostic: per-stage clock64 timeline of CTA 0 of k_mlp_forward (needs a B200).  python tools/mlp_timeline.py [n_envs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from locomanipulationrl_b200 import _lib  # noqa: E402
from locomanipulationrl_b200.policy.mlp import MlpActorCritic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
m = MlpActorCritic(device="cuda:0")
obs = torch.randn(n, 64, device="cuda")
torch.set_grad_enabled(False)          # the inference kernel (the differentiable forward carries no stamps)
for _ in range(3):
    m(obs)
stamps = torch.zeros(256, dtype=torch.int64, device="cuda")
lib = _lib.load()
_lib.check(lib.lm_mlp_set_timeline(m._handle, stamps.data_ptr()), "set_timeline")
m(obs)
torch.cuda.synchronize()
_lib.check(lib.lm_mlp_set_timeline(m._handle, None), "set_timeline")
s = stamps.cpu().tolist()
t0 = s[0]
print(f"weights arrived (control) +{s[1] - t0}; epilogue warps past first A0 +{s[2] - t0}; kernel end +{s[3] - t0}")
names = ["wait acc1", "E1 start", "E1 done", "obs->A0 done", "heads done", "wait acc2", "E2 start", "E2 done"]
it = 0
while s[8 + 8 * it] != 0:
    row = s[8 + 8 * it: 16 + 8 * it]
    print(f"tile {it}: " + "  ".join(f"{nm} +{v - t0}" for nm, v in zip(names, row)))
    it += 1
