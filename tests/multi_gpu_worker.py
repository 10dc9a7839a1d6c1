"""Worker of tests/test_multi_gpu.py (run under torchrun, one rank per GPU).  Checks, with the reference yaml's update settings
(5 mini-epochs, adaptive legacy schedule at kl_threshold 0.008):
  1. PeerComm.all_reduce (one-shot NVLink peer-memory kernel) == torch.distributed.all_reduce, eager and from a CUDA graph;
  2. PpoUpdater over different rollouts per rank leaves bit-identical parameters AND learning rates on every rank — through NCCL and
     through the all-reduce fused into the weight-gradient reduction — and the two transports agree;
  3. the sharded update differs from a rank updating alone (the collective really mixes the ranks).
Exit code 0 = all checks passed on this rank."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from locomanipulationrl_b200.dist import PeerComm  # noqa: E402
from locomanipulationrl_b200.policy.mlp import MlpActorCritic  # noqa: E402
from locomanipulationrl_b200.ppo import PpoUpdater  # noqa: E402
from locomanipulationrl_b200.rollout import RolloutBuffer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
T, N, MB = 4, 2048, 2048
fails = []


def check(ok, what):
    if not ok:
        fails.append(what)
    print(f"rank {rank}: {'ok  ' if ok else 'FAIL'} {what}", flush=True)


# ---- 1. the stand-alone collective
comm = PeerComm(max_floats=70000, device=dev)
for n in (1, 14, 4099, 58651):
    g = torch.Generator(device=dev).manual_seed(1000 * rank + n)
    x = torch.randn(n, device=dev, generator=g)
    ref = x.clone()
    dist.all_reduce(ref)
    for rep in range(3):                                    # both parities of the receive slots, back to back
        y = x.clone()
        comm.all_reduce(y)
        tol = 0.0 if world == 2 else 1e-5                   # two addends commute exactly; more ranks: NCCL's own summation order
        check(float((y - ref).abs().max()) <= tol * float(ref.abs().max() + 1), f"PeerComm.all_reduce n={n} rep={rep}")
    gathered = [torch.empty_like(y) for _ in range(world)]
    dist.all_gather(gathered, y)
    check(all(torch.equal(gathered[0], t) for t in gathered), f"PeerComm result bit-identical on every rank n={n}")
x = torch.randn(58651, device=dev)
ref = x.clone(); dist.all_reduce(ref); ref2 = ref.clone(); dist.all_reduce(ref2)
y = x.clone()
comm.all_reduce(y); y.copy_(x); torch.cuda.synchronize()     # warm
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    comm.all_reduce(y)
gr.replay(); torch.cuda.synchronize()
check(float((y - ref).abs().max()) <= 1e-5 * float(ref.abs().max()), "graph-captured all-reduce, first replay")
gr.replay(); torch.cuda.synchronize()                        # second replay reduces the reduced tensor again
check(float((y - ref2).abs().max()) <= 1e-5 * float(ref2.abs().max()), "graph-captured all-reduce, second replay")
check(comm.status() == 0, "no time-out recorded")

# ---- 1b. the rollout-boundary reducer: statistics gathered into the message by the all-reduce kernel == the NCCL path
from locomanipulationrl_b200.dist import FusedHorizonReducer  # noqa: E402
extras = torch.arange(16, device=dev, dtype=torch.float32) * (rank + 1) * 0.37
counters = torch.arange(1, 7, device=dev) * (3 * rank + 1)
outs = []
for c in (None, comm):
    net = torch.nn.Linear(7, 3).to(dev)
    fr = FusedHorizonReducer(net.parameters(), 100 * (rank + 1), comm=c)
    for p_ in net.parameters():
        p_.grad.fill_(float(2 * rank + 1))
    for rep in range(2):
        fr.reduce(extras, counters)
        if rep == 0:
            for p_ in net.parameters():
                p_.grad.fill_(float(2 * rank + 1))
    torch.cuda.synchronize()
    r = fr.result()
    outs.append((fr.arena.flat.clone() * fr.grad_scale, r["reward_means"].clone(), r["counters"].clone(), int(r["num_envs"])))
check(torch.allclose(outs[0][0], outs[1][0]) and torch.allclose(outs[0][1], outs[1][1], rtol=1e-6) and torch.equal(outs[0][2], outs[1][2])
      and outs[0][3] == outs[1][3] == 100 * world * (world + 1) // 2, "FusedHorizonReducer: peer gather path == NCCL path")


# ---- 2. the update
def make(seed_model, seed_data, comm_=None, **kw):
    torch.manual_seed(seed_model)
    model = MlpActorCritic(device=dev)
    buf = RolloutBuffer(T, N, device=dev)
    up = PpoUpdater(model, buf, minibatch_size=MB, mini_epochs=5, lr=3e-3, comm=comm_, normalize_value=True, **kw)
    g = torch.Generator(device=dev).manual_seed(seed_data)
    ls = model.log_std_parameter.detach().clone()
    with torch.no_grad():
        for t in range(T + 1):
            obs = torch.randn(N, 64, device=dev, generator=g).clamp(-5, 5)
            mu, value, _ = model.reference_forward(obs)
            rew, dn = torch.randn(N, device=dev, generator=g), (torch.rand(N, device=dev, generator=g) < 0.02).long()
            if t == T:
                lv, ld = buf.finish(rew, value.contiguous(), dn, value_denorm=up.value_denorm)
            else:
                buf.write(t, obs, mu.contiguous(), value.contiguous(), ls, rew, dn, noise=torch.randn(N, 12, device=dev, generator=g),
                          value_denorm=up.value_denorm)
    return model, up, lv.clone(), ld.clone()


def same_everywhere(t):
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t.contiguous())
    return all(torch.equal(gathered[0], u) for u in gathered)


results = {}
for name, c in (("nccl", None), ("peer", comm)):
    model, up, lv, ld = make(0, 100 + rank, c)
    up.update(lv, ld)
    torch.cuda.synchronize()
    check(same_everywhere(up.flat), f"{name}: parameters bit-identical on every rank after 5 mini-epochs with the adaptive schedule")
    check(same_everywhere(up.opt.lr_dev), f"{name}: learning rate identical on every rank ({up.opt.lr:.3e})")
    # the schedule itself, deterministically: with a threshold far below any KL the policy shows after its first optimiser step,
    # every later minibatch must shrink the rate (with kl_threshold 0.008 above the rate may legitimately end where it started)
    _, up_s, lv_s, ld_s = make(0, 100 + rank, c, kl_threshold=1e-7)
    up_s.update(lv_s, ld_s)
    check(up_s.opt.lr < 3e-3 * 0.7 and same_everywhere(up_s.opt.lr_dev), f"{name}: the adaptive schedule acts, identically on every rank (lr 3.0e-03 -> {up_s.opt.lr:.3e})")
    probe = torch.randn(512, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(5)).clamp(-5, 5)
    with torch.no_grad():
        mu_k, mu_m = model(probe)[0], model.reference_forward(probe)[0]
    err, mag = float((mu_k - mu_m).abs().max()), float(mu_m.abs().max())
    check(err < 1e-2 * max(1.0, mag), f"{name}: the kernels act with the updated weights (|mu| up to {mag:.2f}, max err {err:.2e})")
    results[name] = (up.flat.clone(), up.opt.lr)
d = float((results["nccl"][0] - results["peer"][0]).abs().max())
check(d <= (0.0 if world == 2 else 1e-4), f"fused NVLink all-reduce == NCCL all-reduce (max |delta| {d:.2e})")
check(comm.status() == 0, "no time-out recorded after the updates")

# ---- 3. a rank on its own
dist.barrier()
mine = results["nccl"][0]
dist.destroy_process_group()
model2, up2, lv2, ld2 = make(0, 100 + rank, None)
up2.update(lv2, ld2)
check(not torch.equal(up2.flat, mine), "the sharded update differs from the single-rank update")
sys.exit(1 if fails else 0)
