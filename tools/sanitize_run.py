"""Small end-to-end exercise of the round-2 kernels (written for compute-sanitizer memcheck, which this pool has closed; it still serves
as a quick plain-run check): default pipeline (action sink + early reset + deferred
finish) for a few eager steps at a ragged size, the host-buffer step with 2 shards, and one PPO minibatch.
    compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from locomanipulationrl_b200 import configs, synthetic  # noqa: E402
from locomanipulationrl_b200.host_step import ShardedHostStepper  # noqa: E402
from locomanipulationrl_b200.policy.mlp import MlpActorCritic  # noqa: E402
from locomanipulationrl_b200.tasks.base import LocoManipulationTask  # noqa: E402

dev = torch.device("cuda:0")
cfg = configs.get_task_config("JointLocomanipulationVertical")
n = 2 * 1153                                   # ragged tiles in both halves
torch.manual_seed(0)
policy = MlpActorCritic(device=dev)
scenes = [synthetic.SyntheticScene(cfg, n, device=dev, seed=5) for _ in range(2)]
for i, sc in enumerate(scenes):
    sc.load(sc.generate(i))
task = LocoManipulationTask(cfg, n, scene=scenes[0], device=dev, fuse_obs_clamp=True)
task.set_defer_finish(True)
task.set_early_reset(True)
mu, value = torch.zeros(n, 12, device=dev), torch.zeros(n, 1, device=dev)
uni = torch.rand(n, 3, device=dev)
task.pre_physics_step(mu, uni)
with torch.no_grad():
    for t in range(6):
        task.set_up_scene(scenes[t % 2])
        task.post_physics_step()
        policy(task.obs_buf, out=(mu, value), sink=task.action_sink())
        task.pre_physics_step(None, uni)
    task.post_finish()
torch.cuda.synchronize()
print("pipeline ok", float(task.rew_buf.sum()), int(task.num_reset_buf[0]))
st = ShardedHostStepper(cfg, 2048, policy, shards=2, device=dev, readback_obs=True)
hs = [s.scene.pack_host(s.scene.generate(0)) for s in st.shards]
for _ in range(3):
    st.step(hs)
torch.cuda.synchronize()
print("host stepper ok", float(st.shards[0].rew_host.sum()))
obs = torch.randn(1000, 64, device=dev)
m, v, _ = policy(obs)                          # training forward + backward on the tcgen05 kernels
(m.sum() + v.sum()).backward()
torch.cuda.synchronize()
print("train forward / backward ok", float(policy.net[0].weight.grad.abs().sum()))
