"""Diagnostic: per-stage clock64 timeline of CTA 0 of k_gnn_forward_tc (needs a B200).  python tools/gnn_timeline.py [n_envs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from locomanipulationrl_b200 import _lib  # noqa: E402
from locomanipulationrl_b200.policy.gnn import GnnActorCritic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
m = GnnActorCritic(device="cuda:0", precision="bf16")
obs = torch.randn(n, 64, device="cuda")
for _ in range(3):
    m(obs)
stamps = torch.zeros(512, dtype=torch.int64, device="cuda")
lib = _lib.load()
_lib.check(lib.lm_gnn_set_timeline(m._handle, stamps.data_ptr()), "set_timeline")
m(obs)
torch.cuda.synchronize()
_lib.check(lib.lm_gnn_set_timeline(m._handle, None), "set_timeline")
s = stamps.cpu().tolist()
t0 = s[0]
print(f"kernel end +{s[1] - t0}")
stages = ["E1a", "E1b", "E2a", "E2b", "E1c", "E1d", "E2c", "E2d"]
it = 0
while it < 7 and s[8 + 64 * it] != 0:
    b = 8 + 64 * it
    print(f"tile {it}: start +{s[b] - t0}  input layer done +{s[b + 1] - t0}  layers done +{s[b + 2] - t0} +{s[b + 3] - t0} +{s[b + 4] - t0}  heads done +{s[b + 5] - t0}")
    print(f"   control thread, layer 0: stage-1 issue of leg 1 (24 MMAs) started +{s[b + 6] - t0}, returned after {s[b + 7] - s[b + 6]} cycles")
    e1 = s[b + 8 + 16]          # layer 1, E1a: barrier wait satisfied
    print(f"   layer 1 E1a inside: first 3 TMEM loads landed +{s[b + 56] - e1}, 3 chunks stored +{s[b + 57] - e1}, next 3 loads landed +{s[b + 58] - e1}, "
          f"3 chunks stored +{s[b + 59] - e1}, fence + arrive done +{s[b + 8 + 16 + 1] - e1}")
    for layer in range(3):
        lb = b + 8 + 16 * layer
        prev = s[b + 1 + layer]
        parts = []
        for k, nm in enumerate(stages):
            w, d = s[lb + 2 * k], s[lb + 2 * k + 1]
            parts.append(f"{nm} wait {w - prev} run {d - w}")
            prev = d
        print(f"   layer {layer}: " + " | ".join(parts))
    it += 1
