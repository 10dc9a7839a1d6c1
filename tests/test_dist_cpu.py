"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: env sharding permutation, stat reduction, gradient arena."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from locomanipulationrl_b200.dist import EnvSharding, FusedHorizonReducer, GradArena, StatReducer


def test_sharding_permutation_covers_reference_order():
    n, world = 64, 4
    ids = torch.cat([EnvSharding(n, world, r).global_ids() for r in range(world)])
    assert sorted(ids.tolist()) == list(range(n))
    s1 = EnvSharding(n, world, 1)
    g = s1.global_ids()
    assert g[:8].tolist() == list(range(8, 16)) and g[8:].tolist() == list(range(40, 48))   # local loco half, local mani half
    assert EnvSharding(n, 2, 0, cotrain=False).global_ids().tolist() == list(range(32))
    with pytest.raises(ValueError):
        EnvSharding(30, 4, 0)
    x = torch.arange(n * 2).view(n, 2)
    assert torch.equal(s1.scatter_global(x), x[g])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from locomanipulationrl_b200 import dist as lmdist
    r, w, _ = lmdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # stats: per-rank means weighted by env count, counters summed
    n_local = 100 * (rank + 1)
    extras = torch.zeros(16)
    extras[:7] = float(rank + 1)
    counters = torch.tensor([1, 2, 3, 4, 5, 6]) * (rank + 1)
    out = StatReducer(n_local, "cpu").reduce(extras, counters)
    expect_mean = (1.0 * 100 + 2.0 * 200) / 300
    ok = torch.allclose(out["reward_means"], torch.full((7,), expect_mean)) and out["counters"].tolist() == [3, 6, 9, 12, 15, 18]
    ok &= int(out["num_envs"]) == 300
    # gradient arena: grads are views of one flat buffer; mean all-reduce
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    arena = GradArena(net.parameters())
    assert arena.numel == sum(p.numel() for p in net.parameters())
    for p in net.parameters():
        p.grad.fill_(float(rank + 1))
    assert float(arena.flat.min()) == rank + 1          # .grad aliases the arena
    arena.allreduce_mean()
    ok &= torch.allclose(arena.flat, torch.full_like(arena.flat, 1.5))
    ok &= all(torch.allclose(p.grad, torch.full_like(p.grad, 1.5)) for p in net.parameters())
    # fused rollout-boundary reduction: gradients + stats in one all-reduce
    net2 = torch.nn.Linear(5, 2)
    fr = FusedHorizonReducer(net2.parameters(), n_local)
    for p in net2.parameters():
        p.grad.fill_(float(2 * rank + 1))
    fr.reduce(extras, counters)
    out2 = fr.result()
    ok &= torch.allclose(fr.arena.flat, torch.full_like(fr.arena.flat, 2.0))            # mean of 1 and 3
    ok &= torch.allclose(out2["reward_means"], out["reward_means"], rtol=1e-6) and torch.equal(out2["counters"], out["counters"])
    ok &= int(out2["num_envs"]) == int(out["num_envs"])
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_reducers_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_adapters_construct_without_gpu():
    """Reference-signature task classes parse sim_config like the reference constructors (QPC:108-115)."""
    from locomanipulationrl_b200.tasks.adapters import TASK_MAP, config_from_sim_config

    class Sim:
        config = {"test": False, "sim_device": "cuda:0", "rl_device": "cuda:0",
                  "task": {"env": {"clipObservations": 5.0, "clipActions": 1.0, "controlFrequencyInv": 2}}}
        task_config = {"env": {"numEnvs": 4096, "envSpacing": 0.5}, "sim": {"dt": 0.01, "max_episode_length": 200}}
    cfg, n = config_from_sim_config("JointLocomanipulationVertical", Sim)
    assert n == 4096 and cfg.dt == 0.01 and cfg.control_frequency_inv == 2 and cfg.max_episode_length == 200
    assert abs(cfg.time_interval - 0.02) < 1e-12
    assert set(TASK_MAP) >= {"QuadrupedPoseControl", "QuadrupedManipulatePlate", "JointLocomanipulation", "JointLocomanipulationVertical"}
    from locomanipulationrl_b200.policy import adapters
    b = adapters.RlGamesMlpBuilder()
    b.load({"device": "cpu"})
    net = b.build("locomani_mlp", actions_num=12, input_shape=(64,))
    assert not net.is_rnn() and sum(p.numel() for p in net.parameters()) == 58637 + 12     # SURVEY §2.1 count + log_std (12)
    g = adapters.RlGamesGnnBuilder(); g.load({"device": "cpu"})
    assert sum(p.numel() for p in g.build("locomani_gnn").parameters()) == 10178 + 12  # + log_std (12)
