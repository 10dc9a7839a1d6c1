"""Device time of one env step (pre + post [+ PD loop / add-ons]) for every task file at a few sizes, plus the reset-storm pre step
and the GAE kernels: a sweep for slow paths outside the headline workload.  python tools/kernel_zoo.py  (needs a B200)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from locomanipulationrl_b200 import configs, gae, synthetic  # noqa: E402
from locomanipulationrl_b200.tasks.base import LocoManipulationTask  # noqa: E402

dev = "cuda:0"


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3      # us


for name in sorted(configs.TASK_CONFIGS):
    cfg = configs.get_task_config(name)
    for n in (4096, 65536):
        scene = synthetic.SyntheticScene(cfg, n, device=dev, seed=3)
        scene.load(scene.generate(0))
        task = LocoManipulationTask(cfg, n, scene=scene, device=dev)
        g = torch.Generator(device=dev).manual_seed(1)
        act = torch.rand(n, 12, device=dev, generator=g) * 2 - 1
        uni = torch.rand(n, 3, device=dev, generator=g)
        pd = cfg.control_mode == configs.CONTROL_PD
        sub = (lambda i: None) if pd else None

        def step():
            task.pre_physics_step(act, uni, sub) if pd else task.pre_physics_step(act, uni)
            task.post_physics_step()

        def pre_only():
            task.pre_physics_step(act, uni, sub) if pd else task.pre_physics_step(act, uni)

        def post_only():
            task._posted_since_pre = False
            task.post_physics_step()

        def storm():
            task.reset()
            task.pre_physics_step(act, uni, sub) if pd else task.pre_physics_step(act, uni)

        step()
        row = {"task": name, "envs": n, "step_us": round(timed(step), 1), "pre_us": round(timed(pre_only), 1), "post_us": round(timed(post_only), 1),
               "reset_storm_pre_us": round(timed(storm), 1)}
        print(json.dumps(row), flush=True)
        del task, scene

T, N = 16, 65536
r, v = torch.randn(T, N, device=dev), torch.randn(T, N, device=dev)
d = (torch.rand(T, N, device=dev) < 0.03).float()
lv, ld = torch.randn(N, device=dev), torch.zeros(N, device=dev)
print(json.dumps({"kernel": "gae time-major [16,65536]", "us": round(timed(lambda: gae.compute_gae(r, v, d, lv, ld)), 1)}))
rt, vt, dt = r.t().contiguous(), v.t().contiguous(), d.t().contiguous()
print(json.dumps({"kernel": "gae env-major [65536,16]", "us": round(timed(lambda: gae.compute_gae(rt, vt, dt, lv, ld, env_major=True)), 1)}))
