"""Diagnostic: per-CTA stage timeline (%globaltimer) of k_post_pair (needs a B200).  python tools/post_timeline.py [n_envs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from locomanipulationrl_b200 import _lib, configs, synthetic  # noqa: E402
from locomanipulationrl_b200.tasks.base import LocoManipulationTask  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda:0")
cfg = configs.get_task_config("JointLocomanipulationVertical")
scenes = [synthetic.SyntheticScene(cfg, n, device=dev, seed=42) for _ in range(6)]
for i, sc in enumerate(scenes):
    sc.load(sc.generate(i))
task = LocoManipulationTask(cfg, n, scene=scenes[0], device=dev, fuse_obs_clamp=True)
mu = torch.zeros(n, 12, device=dev)
uni = torch.rand(n, 3, device=dev)
for i in range(6):
    task.pre_physics_step(mu, uni)
    task.set_up_scene(scenes[i])
    task.post_physics_step()
tiles = 2 * ((n // 2 + 127) // 128)
stamps = torch.zeros(tiles, 8, dtype=torch.int64, device=dev)
lib = _lib.load()
_lib.check(lib.lm_task_set_timeline(task._handle, stamps.data_ptr()), "set_timeline")
task.pre_physics_step(mu, uni)
task.set_up_scene(scenes[0])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
task.post_physics_step()
e1.record()
torch.cuda.synchronize()
_lib.check(lib.lm_task_set_timeline(task._handle, None), "set_timeline")
s = stamps.cpu().double()
t0 = s[:, 0].min()
s = (s - t0) / 1e3
names = ["start", "loads issued", "TMA landed", "math+stores issued", "finish done"]
print(f"n={n} tiles={tiles} event time {e0.elapsed_time(e1) * 1e3:.1f} us")
for k, nm in enumerate(names):
    c = s[:, k]
    print(f"{nm:22s} min {c.min():7.2f}  p50 {c.median():7.2f}  p90 {c.quantile(0.9):7.2f}  max {c.max():7.2f} us")
d = s[:, 3] - s[:, 2]
print(f"math phase per CTA: p50 {d.median():.2f} max {d.max():.2f} us;  wait for TMA after loads issued: p50 {(s[:,2]-s[:,1]).median():.2f} us")
