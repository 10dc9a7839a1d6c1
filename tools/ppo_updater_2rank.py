"""2-rank check of PpoUpdater's gradient all-reduce (run under torchrun with 2 GPUs):
same initial weights, a different synthetic rollout per rank -> after one update the parameters must be bit-identical on both ranks
and must differ from the update a rank computes on its own rollout alone."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from locomanipulationrl_b200.policy.mlp import MlpActorCritic  # noqa: E402
from locomanipulationrl_b200.ppo import PpoUpdater  # noqa: E402
from locomanipulationrl_b200.rollout import RolloutBuffer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
T, N, MB = 4, 1024, 2048


def make(seed_model, seed_data):
    torch.manual_seed(seed_model)
    model = MlpActorCritic(device=dev)
    buf = RolloutBuffer(T, N, device=dev)
    g = torch.Generator(device=dev).manual_seed(seed_data)
    ls = model.log_std_parameter.detach().clone()
    with torch.no_grad():
        for t in range(T):
            obs = torch.randn(N, 64, device=dev, generator=g).clamp(-5, 5)
            mu, value, _ = model.reference_forward(obs)
            buf.write(t, obs, mu.contiguous(), value.contiguous(), ls, torch.randn(N, device=dev, generator=g),
                      torch.zeros(N, dtype=torch.long, device=dev), noise=torch.randn(N, 12, device=dev, generator=g))
    return model, buf, torch.randn(N, device=dev, generator=g), torch.zeros(N, device=dev)


model, buf, lv, ld = make(0, 100 + rank)
up = PpoUpdater(model, buf, minibatch_size=MB, mini_epochs=1, kl_threshold=None)
up.update(lv, ld)
mine = up.flat.clone()
gathered = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(gathered, mine)
same = all(torch.equal(gathered[0], g) for g in gathered)

# the same rank alone (no process group visible to the updater): must give a different result
dist.barrier()
dist.destroy_process_group()
model2, buf2, lv2, ld2 = make(0, 100 + rank)
up2 = PpoUpdater(model2, buf2, minibatch_size=MB, mini_epochs=1, kl_threshold=None)
up2.update(lv2, ld2)
differs = not torch.equal(up2.flat, mine)
print(f"rank {rank}: identical across ranks = {same}; differs from the single-rank update = {differs}; max |delta| = {float((up2.flat - mine).abs().max()):.3e}", flush=True)
sys.exit(0 if (same and differs) else 1)
