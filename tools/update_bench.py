"""Device times of the caller-side kernels (rollout writer, running statistics, noise, PPO loss head, Adam) at the bench sizes.
python tools/update_bench.py  (needs a B200; CUDA events around graph replays, inputs rotated over > L2 worth of buffers)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from locomanipulationrl_b200.noise import NoiseRandomizer  # noqa: E402
from locomanipulationrl_b200.ppo import FlatAdam, PpoLoss  # noqa: E402
from locomanipulationrl_b200.rollout import RolloutBuffer, RunningMeanStd  # noqa: E402

dev = "cuda:0"
N, T, SETS = 65536, 16, 8
HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def graph_time(fn, calls, reps=20):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(calls):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * calls)


def report(name, ms, bytes_per_item, items):
    gbs = items * bytes_per_item / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "us": round(ms * 1e3, 2), "items": items, "bytes_per_item": bytes_per_item, "GB/s": round(gbs, 1), "frac_of_measured_hbm": round(gbs / HBM, 3)}))


g = torch.Generator(device=dev).manual_seed(0)
obs = [torch.randn(N, 64, device=dev, generator=g) for _ in range(SETS)]          # 8 x 16.8 MB > L2
mu = [torch.randn(N, 12, device=dev, generator=g) for _ in range(SETS)]
noise = [torch.randn(N, 12, device=dev, generator=g) for _ in range(SETS)]
value = torch.randn(N, 1, device=dev, generator=g)
rew = torch.randn(N, device=dev, generator=g)
dones = (torch.rand(N, device=dev, generator=g) < 0.03).long()
log_std = torch.zeros(12, device=dev)
actions = torch.zeros(N, 12, device=dev)
buf, stats = RolloutBuffer(T, N, device=dev, reward_scale=0.01), RunningMeanStd(64, device=dev)
for i in range(3):
    buf.write(i, obs[i], mu[i], value, log_std, rew, dones, noise=noise[i], actions_out=actions, obs_stats=stats)
torch.cuda.synchronize()
ms = graph_time(lambda i: buf.write(i % T, obs[i % SETS], mu[i % SETS], value, log_std, rew, dones, noise=noise[i % SETS], actions_out=actions, obs_stats=stats), T)
report("k_rollout_write (65,536 envs, sampled actions, fused obs statistics)", ms, 368 + 48 + 368 + 48, N)

vals = torch.randn(T * N, device=dev, generator=g)
vstats = RunningMeanStd(1, device=dev)
vstats.update(vals)
ms = graph_time(lambda i: vstats.update(vals), 4)
report("k_rms_update (value statistics, 16 x 65,536 samples)", ms, 4, T * N)

rz = NoiseRandomizer({"on_reset": {"operation": "additive", "distribution": "gaussian", "distribution_parameters": [0.0, 0.0001]},
                      "on_interval": {"frequency_interval": 1, "operation": "additive", "distribution": "gaussian", "distribution_parameters": [0.0, 0.002]}},
                     N, 64, device=dev, seed=1)
rz.apply(obs[0], dones)
ms = graph_time(lambda i: rz.apply(obs[i % SETS], dones), SETS)
report("k_noise_apply (observations, in-kernel Philox draws)", ms, 256 * 3, N)

B = T * N // 4                                                                     # rl_games minibatch: a quarter of the rollout
pm, pa, pom = (torch.randn(B, 12, device=dev, generator=g) * 0.3 for _ in range(3))
pv, pnl, padv, pov, pret = (torch.randn(B, device=dev, generator=g) for _ in range(5))
head = PpoLoss(device=dev)
out = {"grad_mu": torch.empty(B, 12, device=dev), "grad_value": torch.empty(B, device=dev)}
head(pm, pv, log_std, pa, pnl, padv, pov, pret, pom, log_std, out=out)
# three rotating input sets (3 x 57 MB > L2): replaying one set alone would time L2 hits, not HBM
psets = [(pm, pv, pa, pnl, padv, pov, pret, pom)] + [tuple(t.clone() for t in (pm, pv, pa, pnl, padv, pov, pret, pom)) for _ in range(2)]


def loss_call(i):
    m_, v_, a_, nl_, ad_, ov_, rt_, om_ = psets[i % 3]
    head(m_, v_, log_std, a_, nl_, ad_, ov_, rt_, om_, log_std, out=out)


ms = graph_time(loss_call, 6)
report(f"k_ppo_loss (minibatch {B}, inputs rotated past L2)", ms, 3 * 48 + 5 * 4 + 48 + 4, B)

flat, grads = torch.randn(58637, device=dev, generator=g), torch.randn(58637, device=dev, generator=g)
opt = FlatAdam(flat)
opt.step(grads)
ms = graph_time(lambda i: opt.step(grads), 8)
report("k_adam_step (58,637 parameters, clip + Adam)", ms, 28, 58637)

# ---------------------------------------------------------------- the MLP update on tcgen05 vs torch autograd on the same GPU
from locomanipulationrl_b200 import _lib  # noqa: E402
from locomanipulationrl_b200.policy.mlp import MlpActorCritic  # noqa: E402
from locomanipulationrl_b200.ppo import MlpPpoGradients  # noqa: E402


def torch_ppo_loss(mu, value, log_std, actions, old_neglogp, advantages, old_values, returns, e_clip=0.2, critic_coef=2.0, bounds_loss_coef=1e-4):
    """The rl_games loss in stock torch ops, written out here for the torch-autograd timing baseline only (the checked
    restatement lives with the tests; this tool must not reach into oracle/)."""
    import math
    sigma = torch.exp(log_std)
    nlp = 0.5 * (((actions - mu) / sigma) ** 2).sum(-1) + 0.5 * math.log(2.0 * math.pi) * actions.shape[-1] + log_std.sum(-1)
    ratio = torch.exp(old_neglogp - nlp)
    a_loss = torch.max(-advantages * ratio, -advantages * torch.clamp(ratio, 1.0 - e_clip, 1.0 + e_clip))
    vpc = old_values + (value - old_values).clamp(-e_clip, e_clip)
    c_loss = torch.max((value - returns) ** 2, (vpc - returns) ** 2)
    b_loss = (torch.clamp_max(mu + 1.1, 0.0) ** 2 + torch.clamp_min(mu - 1.1, 0.0) ** 2).sum(-1)
    return a_loss.mean() + 0.5 * c_loss.mean() * critic_coef + b_loss.mean() * bounds_loss_coef


torch.manual_seed(0)
model = MlpActorCritic(device=dev)
model.pack()
pobs = torch.randn(B, 64, device=dev, generator=g).clamp(-5, 5)
grads = MlpPpoGradients(model, max_batch=B)
lib = _lib.load()
cs = lambda: torch.cuda.current_stream().cuda_stream      # graph capture runs on its own stream
p = _lib.ptr


def run_all(i=0):
    grads(pobs, pa, pnl, padv, pov, pret, pom, log_std)


run_all()
torch.cuda.synchronize()
ms_all = graph_time(run_all, 2, reps=10)
ms_fwd = graph_time(lambda i: lib.lm_mlp_forward_train(model._handle, p(pobs), None, None, 5.0, p(grads.mu), p(grads.value), p(grads.a0), p(grads.a1),
                                                       p(grads.a2), p(grads.a3), B, cs()), 2, reps=10)
ms_bwd = graph_time(lambda i: lib.lm_mlp_backward(model._handle, p(grads.grad_mu), p(grads.grad_value), p(grads.a1), p(grads.a2), p(grads.a3), p(grads.dh),
                                                  p(grads.dz3), p(grads.dz2), p(grads.dz1), B, cs()), 2, reps=10)
prm = [model.net[0].weight, model.net[0].bias, model.net[2].weight, model.net[2].bias, model.net[4].weight, model.net[4].bias,
       model.mean_layer.weight, model.mean_layer.bias, model.value_layer.weight, model.value_layer.bias]
ms_wg = graph_time(lambda i: lib.lm_mlp_wgrad(p(grads.dz1), p(grads.dz2), p(grads.dz3), p(grads.dh), p(grads.a0), p(grads.a1), p(grads.a2), p(grads.a3), B,
                                              p(grads._ws), grads._ctas, *[q.grad.data_ptr() for q in prm], cs()), 2, reps=10)
FLOP = 3 * 116352 * B
print(json.dumps({"update": f"MLP PPO gradients on tcgen05, minibatch {B}", "us_total": round(ms_all * 1e3, 1), "us_forward_train": round(ms_fwd * 1e3, 1),
                  "us_backward_act": round(ms_bwd * 1e3, 1), "us_wgrad": round(ms_wg * 1e3, 1), "TFLOP/s (3x forward FLOP)": round(FLOP / (ms_all * 1e-3) / 1e12, 1),
                  "hbm_bytes_per_sample": 64 * 4 + 2 * (64 + 256 + 128 + 64) + 2 * 2 * (256 + 128 + 64) + 2 * 2 * (256 + 128 + 64 + 16) + 2 * (64 + 256 + 128 + 64)}))


def torch_step(autocast):
    for q in model.parameters():
        q.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        mu, value, ls = model.reference_forward(pobs)
    loss = torch_ppo_loss(mu.float(), value.float().view(-1), ls, pa, pnl, padv, pov, pret)
    loss.backward()


for ac in (False, True):
    for _ in range(3):
        torch_step(ac)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch_step(ac)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"baseline": f"torch autograd, same GPU, {'bf16 autocast' if ac else 'fp32 (TF32 off)'}, minibatch {B}", "us_total": round(e0.elapsed_time(e1) / 10 * 1e3, 1)}))
