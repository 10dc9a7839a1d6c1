"""Parity at the shapes and through the entry points bench.py times: the 65,536-env co-train Pipeline under
CUDA-graph replay with the fused observation clamp, the > 1,024-tile finish path (k_post_finish), the host-buffer stepper behind
the e2e figure, shard invariance of the env sharding, and the position / effort branches of take_action."""
import pytest
import torch

from locomanipulationrl_b200 import configs, synthetic
from tests.common import CudaRollout, assert_close, rollout_inputs

pytestmark = pytest.mark.gpu


def _tracker(task, port, scene_cpu, cfg, n):
    """A CudaRollout shell around an existing (task, port) pair: reuses compare_post()."""
    ro = CudaRollout.__new__(CudaRollout)
    ro.cfg, ro.n, ro.task, ro.port, ro.scene_cpu, ro.scene_gpu = cfg, n, task, port, scene_cpu, None
    ro.exempt_events, ro.max_err = 0, {}
    return ro


def test_bench_default_pipeline_65536_graph_replay_matches_port():
    """BASELINE configs[2] exactly as bench.py times it by default: JointLocomanipulationVertical, 65,536 envs, fuse_obs_clamp, six
    rotating state sets, action = policy mean delivered through the action sink (the action half of pre_physics_step inside the MLP head
    epilogue), the reset half of the next pre_physics_step released beside the policy kernel by programmatic dependent launch
    (lm_task_set_early_reset), the post kernel's finish pass deferred into that launch, one CUDA graph per state set — tracked by the
    oracle port for the six eager warm-up steps and eight replayed ones.  One pipeline step = post(t) -> policy(t) || pre(t+1); the
    port runs the same calls in order."""
    import bench
    from oracle import port
    dev = torch.device("cuda:0")
    pipe = bench.Pipeline("cotrain64k", dev, rank=0)
    cfg, n, task = pipe.cfg, pipe.n, pipe.task
    S = bench.STATE_SETS
    assert n == 65536 and task._fuse_obs_clamp and pipe.use_sink and pipe.early_reset and pipe.defer_finish
    scenes_cpu = []
    for i in range(S):
        sc = synthetic.SyntheticScene(cfg, n, seed=42)
        sc.load(sc.generate(i))
        scenes_cpu.append(sc)
    p = port.TaskPort(cfg, scenes_cpu[0])
    unis = [u.cpu() for u in pipe.uniforms]
    p.pre_physics_step(torch.zeros(n, 12), unis[0])                     # the pipeline's priming pre_physics_step
    stats = {"exempt": 0, "err": {}}

    def close(name, got, ref, **kw):
        stats["err"][name] = max(stats["err"].get(name, 0.0), assert_close(name, got, ref, **kw))

    def track(i, t):
        if pipe.defer_finish:
            pipe.task.post_finish()        # extras / counters of post(t) now (the captured pre launches carry the pass too: it acts once)
        torch.cuda.synchronize()
        # ---- post(t)
        p.scene = scenes_cpu[i]
        p.post_physics_step()
        obs_ref = torch.clamp(p.obs_buf, -cfg.clip_observations, cfg.clip_observations)        # _process_data, fused into the kernel
        close(f"step {t} obs", task.obs_buf, obs_ref)
        from tests.common import rot_reward_conditioning
        close(f"step {t} rew", task.rew_buf, p.rew_buf, scale=p.inter["rew_scale"], extra_atol=rot_reward_conditioning(p))
        for k, v in p.extras.items():
            close(f"step {t} extras[{k}]", task.extras[k], v, rtol=2e-5, atol=1e-6)
        c = task._counters.cpu()
        for j, k in enumerate(["num_successes", "num_resets", "num_successes_loco", "num_resets_loco", "num_successes_mani", "num_resets_mani"]):
            assert int(c[j]) == int(p.counters[k]), f"step {t}: counter {k}"
        # ---- policy(t): the tcgen05 forward on the observations just produced
        with torch.no_grad():
            mu_ref, v_ref, _ = pipe.policy.reference_forward((task.obs_buf - pipe.norm_mean) * pipe.norm_inv_std)
        close(f"step {t} mu", pipe.mu, mu_ref, rtol=0, atol=1e-2)
        close(f"step {t} value", pipe.value, v_ref, rtol=0, atol=1e-2)
        # ---- pre(t+1): reset half on the second stream, action half inside the MLP head epilogue
        mu_cpu = pipe.mu.detach().cpu()
        p.pre_physics_step(torch.clamp(mu_cpu, -1.0, 1.0), unis[(i + 1) % S])                 # the adapter's clamp (vec_env_rlgames.py:60)
        assert torch.equal(task.reset_env_ids().cpu(), p.reset_env_ids), f"step {t}: reset id list (= the post step's reset mask)"
        for k in ("reset_buf", "progress_buf", "successes", "consecutive_successes", "goal_reset_buf"):
            assert torch.equal(getattr(task, k).cpu(), getattr(p, k)), f"step {t}: {k} after the reset half"
        close(f"step {t} goals", task.goal_quaternions, p.goal_quaternions)
        assert torch.equal(task.current_actions.cpu(), torch.clamp(mu_cpu, -1.0, 1.0)), f"step {t}: current_actions from the sink"
        close(f"step {t} last_actions", task._actions_ring[bench_parity(pipe) ^ 1], p.last_actions, rtol=0, atol=0)
        close(f"step {t} targets_loco", task.targets_loco, p.targets["loco"])
        close(f"step {t} targets_mani", task.targets_mani, p.targets["mani"])
        close(f"step {t} last_tip", task.last_base_tip_positions, p.last_base_tip_positions)
        close(f"step {t} last_jv", task.last_joint_velocities, torch.cat([p.last_joint_velocities["loco"], p.last_joint_velocities["mani"]]))
        for grp in ("loco", "mani"):                                      # the reset scatter landed in the live (rotating) views
            for k, v in getattr(scenes_cpu[i], grp).items():
                close(f"step {t} scene.{grp}.{k}", getattr(pipe.scenes[i], grp)[k], v)

    def bench_parity(pl):
        return pl.lib.lm_task_step_parity(pl.task._handle)

    pipe.capture(after_step=lambda i: track(i, i))
    assert pipe.graphs is not None and len(pipe.graphs) == S
    for t in range(S, S + 8):
        pipe.step(t)
        track(t % S, t)
    print("65,536-env graph replay max abs err:", {k.split(" ", 2)[-1]: f"{v:.2e}" for k, v in sorted(stats["err"].items())[-12:]})



@pytest.mark.parametrize("variant", ["sink_late_inline", "plain"])
def test_bench_pipeline_65536_graph_replay_matches_port(variant, monkeypatch):
    """bench.py's conservative pipelines (step order pre -> post -> policy) at the same shape: sink_late_inline = action sink only, reset
    half behind the policy kernel, finish pass inside the post kernel (LM_BENCH_NO_EARLY, LM_BENCH_INLINE_FINISH); plain = actions
    through k_pre (LM_BENCH_NO_SINK), deferred finish."""
    import bench
    from oracle import port
    if variant == "sink_late_inline":
        monkeypatch.setenv("LM_BENCH_NO_EARLY", "1")
        monkeypatch.setenv("LM_BENCH_INLINE_FINISH", "1")
    if variant == "plain":
        monkeypatch.setenv("LM_BENCH_NO_SINK", "1")
    dev = torch.device("cuda:0")
    pipe = bench.Pipeline("cotrain64k", dev, rank=0)
    cfg, n = pipe.cfg, pipe.n
    S = bench.STATE_SETS
    assert n == 65536 and pipe.task._fuse_obs_clamp
    assert (pipe.use_sink, pipe.early_reset, pipe.defer_finish) == {"sink_late_inline": (True, False, False), "plain": (False, False, True)}[variant]
    scenes_cpu = []
    for i in range(S):
        sc = synthetic.SyntheticScene(cfg, n, seed=42)
        sc.load(sc.generate(i))
        scenes_cpu.append(sc)
    p = port.TaskPort(cfg, scenes_cpu[0])
    ro = _tracker(pipe.task, p, scenes_cpu[0], cfg, n)
    unis = [u.cpu() for u in pipe.uniforms]
    actions = {"a": pipe.actions.detach().cpu().clone()}

    def track(i, t):
        if pipe.defer_finish:
            pipe.task.post_finish()        # extras / counters of post(t) now (the captured pre launches carry the pass too: it acts once)
        torch.cuda.synchronize()
        p.pre_physics_step(torch.clamp(actions["a"], -1.0, 1.0), unis[i])      # the adapter's clamp (vec_env_rlgames.py:60)
        prev = scenes_cpu[(i - 1) % S] if t > 0 else scenes_cpu[0]             # the views that were live during pre_physics_step
        got_prev = pipe.scenes[(i - 1) % S] if t > 0 else pipe.scenes[0]
        for grp in ("loco", "mani"):
            for k in ("dof_pos", "root_pos" if grp == "loco" else "plate_pos", "indicator_quat"):
                assert_close(f"step {t} scene.{grp}.{k}", getattr(got_prev, grp)[k], getattr(prev, grp)[k])
        p.scene = scenes_cpu[i]
        ro.scene_cpu = scenes_cpu[i]
        p.post_physics_step()
        p.obs_buf = torch.clamp(p.obs_buf, -cfg.clip_observations, cfg.clip_observations)      # _process_data, fused into the kernel
        ro.compare_post(t)
        with torch.no_grad():
            mu_ref, v_ref, _ = pipe.policy.reference_forward((pipe.task.obs_buf - pipe.norm_mean) * pipe.norm_inv_std)
        assert_close(f"step {t} mu", pipe.mu, mu_ref, rtol=0, atol=1e-2)
        assert_close(f"step {t} value", pipe.value, v_ref, rtol=0, atol=1e-2)
        actions["a"] = pipe.mu.detach().cpu().clone()

    pipe.capture(after_step=lambda i: track(i, i))
    assert pipe.graphs is not None and len(pipe.graphs) == S
    for t in range(S, S + 8):
        pipe.step(t)
        track(t % S, t)
    assert ro.exempt_events == 0, f"{ro.exempt_events} near-threshold mask exemptions"
    print(f"65,536-env pipeline ({variant}), graph replay, max abs err:", {k: f"{v:.2e}" for k, v in ro.max_err.items()})


@pytest.mark.parametrize("name,n", [("JointLocomanipulationVertical", 4096), ("QuadrupedManipulatePlate", 1000), ("QuadrupedPoseControl", 130)])
def test_action_sink_equals_plain_pre_step(name, n):
    """The action half of pre_physics_step folded into the producing kernels (MLP head epilogue: action = mean; rollout writer:
    action = the sample) leaves bit-identical current_actions / joint targets / task trajectories to the plain call."""
    from locomanipulationrl_b200.policy.mlp import MlpActorCritic
    from locomanipulationrl_b200.rollout import RolloutBuffer
    from locomanipulationrl_b200.tasks.base import LocoManipulationTask
    cfg = configs.get_task_config(name)
    torch.manual_seed(3)
    policy = MlpActorCritic(device="cuda:0")
    with torch.no_grad():
        policy.mean_layer.weight.mul_(40.0)                            # push a good share of the means beyond the +-1 clamp
    sa, sb = synthetic.SyntheticScene(cfg, n, device="cuda:0", seed=4), synthetic.SyntheticScene(cfg, n, device="cuda:0", seed=4)
    a, b = LocoManipulationTask(cfg, n, scene=sa), LocoManipulationTask(cfg, n, scene=sb)
    buf_a, buf_b = RolloutBuffer(4, n), RolloutBuffer(4, n)
    log_std = torch.full((12,), -0.5, device="cuda")
    _, unis = rollout_inputs(n, 8, seed=6)
    act_a = torch.zeros(n, 12, device="cuda")
    mu_a, v_a, mu_b, v_b = (torch.zeros(n, k, device="cuda") for k in (12, 1, 12, 1))
    for t in range(8):
        a.pre_physics_step(act_a, unis[t].cuda())
        b.pre_physics_step(None, unis[t].cuda())                       # delivered by the previous step's producer (zeros at t = 0)
        assert torch.equal(a.current_actions, b.current_actions), f"step {t}: current_actions"
        assert torch.equal(a.targets_loco, b.targets_loco) and torch.equal(a.targets_mani, b.targets_mani), f"step {t}: targets"
        st = sa.generate(t, a.goal_quaternions)
        sa.load(st); sb.load(st)
        a.post_physics_step(); b.post_physics_step()
        assert torch.equal(a.obs_buf, b.obs_buf) and torch.equal(a.rew_buf, b.rew_buf) and torch.equal(a.reset_buf, b.reset_buf), f"step {t}"
        if t < 4:                                                      # evaluation: action = policy mean, delivered by the head epilogue
            policy(a.obs_buf, out=(mu_a, v_a))
            policy(b.obs_buf, out=(mu_b, v_b), sink=b.action_sink())
            assert torch.equal(mu_a, mu_b) and torch.equal(v_a, v_b)
            act_a = mu_a.clone()
        else:                                                          # training: action = the sample, delivered by the rollout writer
            policy(a.obs_buf, out=(mu_a, v_a)); policy(b.obs_buf, out=(mu_b, v_b))
            noise = torch.randn(n, 12, device="cuda")
            buf_a.write(t - 4, a.obs_buf, mu_a, v_a, log_std, a.rew_buf, a.reset_buf, noise=noise, actions_out=act_a)
            buf_b.write(t - 4, b.obs_buf, mu_b, v_b, log_std, b.rew_buf, b.reset_buf, noise=noise, sink=b.action_sink())
        assert float((act_a.abs() > 1.0).float().mean()) > 0.02       # the clamp is exercised


def test_large_grid_finish_kernel_262144_envs():
    """2,048 tiles: beyond the 1,024-tile in-kernel tail, the finish pass runs as k_post_finish (the path behind bench.py's
    `roofline_large`).  Reset id lists (the scan the finish pass leaves), counters and extras are compared inside step()."""
    n = 262144
    ro = CudaRollout("JointLocomanipulationVertical", n)
    ro.cond_ulps = 10.0        # 786,432 env-steps reach the tail of the |quat_diff.xyz| -> 1 conditioning (see rot_reward_conditioning)
    acts, unis = rollout_inputs(n, 3, seed=21)
    for t in range(3):
        ro.step(t, acts[t], unis[t])
    assert ro.exempt_events <= 2
    print("262,144 envs max abs err:", {k: f"{v:.2e}" for k, v in ro.max_err.items()}, "exempt:", ro.exempt_events)


@pytest.mark.parametrize("shards", [1, 4])
def test_host_stepper_matches_port(shards):
    """ShardedHostStepper / lm_task_step_host_async — the entry behind bench.py's e2e figure — against the port + the fp32 torch MLP:
    rewards, reset flags, observations, policy mean and value as they arrive in the pinned host buffers, actions fed back."""
    from locomanipulationrl_b200.host_step import ShardedHostStepper
    from locomanipulationrl_b200.policy.mlp import MlpActorCritic
    from oracle import port
    cfg = configs.get_task_config("JointLocomanipulationVertical")
    n = 8192
    torch.manual_seed(0)
    policy = MlpActorCritic(device="cuda:0")
    norm_mean, norm_inv = torch.randn(64, device="cuda") * 0.1, torch.rand(64, device="cuda") + 0.5
    st = ShardedHostStepper(cfg, n, policy, shards=shards, device="cuda:0", seed=5, env_id_offset=0, norm_mean=norm_mean, norm_inv_std=norm_inv,
                            readback_obs=True)
    g = torch.Generator().manual_seed(9)
    ports, scenes = [], []
    for k, s in enumerate(st.shards):
        sc = synthetic.SyntheticScene(cfg, s.n, seed=5 + 17 * k, env_id_offset=k * s.n)
        scenes.append(sc)
        ports.append(port.TaskPort(cfg, sc))
    exempt = 0
    for t in range(5):
        for s in st.shards:
            s.uniforms_host.copy_(torch.rand(s.n, 3, generator=g))
        actions_in = [a.clone() for a in st.mu_host]
        states = [sc.generate(t) for sc in scenes]
        st.step([s.scene.pack_host(x) for s, x in zip(st.shards, states)])
        for k, (s, p, sc) in enumerate(zip(st.shards, ports, scenes)):
            p.pre_physics_step(torch.clamp(actions_in[k], -1.0, 1.0), s.uniforms_host.clone())
            sc.load(states[k])
            p.post_physics_step()
            ro = _tracker(s.task, p, sc, cfg, s.n)
            p.obs_buf = torch.clamp(p.obs_buf, -5.0, 5.0)
            ro.compare_post(t)
            exempt += ro.exempt_events
            assert_close(f"step {t} shard {k} obs_host", s.obs_host, p.obs_buf)
            assert torch.equal(s.reset_host, s.task.reset_buf.cpu()) and torch.equal(s.rew_host, s.task.rew_buf.cpu())
            with torch.no_grad():
                mu_ref, v_ref, _ = policy.reference_forward(torch.clamp((s.task.obs_buf - norm_mean) * norm_inv, -5.0, 5.0))
            assert_close(f"step {t} shard {k} mu_host", st.mu_host[k], mu_ref.cpu(), rtol=0, atol=1e-2)
            assert_close(f"step {t} shard {k} value_host", s.value_host, v_ref.cpu(), rtol=0, atol=1e-2)
    assert exempt <= 2


def test_shard_invariance_two_shards_equal_one_task():
    """N envs on one rank == N/k envs on each of k ranks after the fixed permutation EnvSharding.global_ids.  Two
    half-size tasks fed the slices of the global state / actions (and the uniform rows their reset ranks map to) reproduce the
    global task bit for bit; counters add up."""
    from locomanipulationrl_b200.dist import EnvSharding
    from locomanipulationrl_b200.tasks.base import LocoManipulationTask
    cfg = configs.get_task_config("JointLocomanipulationVertical")
    N, K = 4096, 2
    H = N // 2
    glob_scene = synthetic.SyntheticScene(cfg, N, device="cuda:0", seed=31)
    glob = LocoManipulationTask(cfg, N, scene=glob_scene)
    shards = []
    for r in range(K):
        sh = EnvSharding(N, K, r)
        sc = synthetic.SyntheticScene(cfg, sh.local_envs, device="cuda:0", seed=31)
        hl = sh.local_envs // 2
        sc.env_pos.copy_(glob_scene.env_pos[r * hl:(r + 1) * hl])
        shards.append((sh, sc, LocoManipulationTask(cfg, sh.local_envs, scene=sc), sh.global_ids().cuda()))
    acts, unis = rollout_inputs(N, 10, seed=13)
    reset_prev = torch.ones(N, dtype=torch.long)
    for t in range(10):
        ids_g = reset_prev.nonzero().squeeze(-1)                     # the global reset list: uniform row = position in it
        rank_of = torch.full((N,), -1, dtype=torch.long)
        rank_of[ids_g] = torch.arange(ids_g.numel())
        glob.pre_physics_step(acts[t].cuda(), unis[t].cuda())
        for sh, sc, task, gids in shards:
            local_reset = reset_prev[gids.cpu()].nonzero().squeeze(-1)
            u = torch.zeros(sh.local_envs, 3)
            u[:local_reset.numel()] = unis[t][rank_of[gids.cpu()[local_reset]]]
            task.pre_physics_step(acts[t][gids.cpu()].cuda(), u.cuda())
        st = glob_scene.generate(t, glob.goal_quaternions)
        glob_scene.load(st)
        for r, (sh, sc, task, gids) in enumerate(shards):
            hl = sh.local_envs // 2
            part = {}
            for grp in ("loco", "mani"):
                part[grp] = {}
                for k, v in st[grp].items():
                    per = v.shape[0] // H                              # rows per env (tips 4, knees 8)
                    part[grp][k] = v[r * hl * per:(r + 1) * hl * per]
            sc.load(part)
        glob.post_physics_step()
        for sh, sc, task, gids in shards:
            task.post_physics_step()
            for name in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "successes", "consecutive_successes", "goal_reset_buf",
                         "goal_quaternions", "last_joint_velocities"):
                assert torch.equal(getattr(task, name), getattr(glob, name)[gids]), f"step {t}: {name}"
        assert torch.equal(sum(task._counters for _, _, task, _ in shards), glob._counters), f"step {t}: counters"
        ex = sum(task._extras_buf[:7].double() * sh.local_envs for sh, _, task, _ in shards) / N
        assert_close(f"step {t} reward-term means", ex, glob._extras_buf[:7], rtol=1e-5, atol=1e-7)
        reset_prev = glob.reset_buf.cpu().clone()


@pytest.mark.parametrize("mode", [configs.CONTROL_POSITION, configs.CONTROL_EFFORT])
def test_take_action_position_and_effort_modes(mode):
    """robot.py:444-461: position targets (+-pi) and effort targets (+-1.5 N m scattered into the 20-wide DOF row).  Every shipped
    robot config uses velocity mode; these branches exist in the reference and in k_pre."""
    from locomanipulationrl_b200.tasks.base import LocoManipulationTask
    from oracle import port
    cfg = configs.get_task_config("QuadrupedPoseControl").with_(control_mode=mode)
    n = 1000
    scene_gpu, scene_cpu = synthetic.SyntheticScene(cfg, n, device="cuda:0", seed=3), synthetic.SyntheticScene(cfg, n, seed=3)
    task, p = LocoManipulationTask(cfg, n, scene=scene_gpu), port.TaskPort(cfg, scene_cpu)
    acts, unis = rollout_inputs(n, 3, seed=8)
    for t in range(3):
        a = acts[t] * 1.7                                             # beyond +-1: the adapter clamp is part of the entry
        p.pre_physics_step(torch.clamp(a, -1.0, 1.0), unis[t])
        task.pre_physics_step(a.cuda(), unis[t].cuda())
        ref = p.targets["loco"]
        if mode == configs.CONTROL_EFFORT:
            assert task.targets_loco.shape == (n, 20)
            full = torch.zeros(n, 20)
            full[:, :12] = ref                                        # dof_index is the identity for the synthetic articulation
            assert_close(f"step {t} effort targets", task.targets_loco, full)
            assert float(ref.abs().max()) <= 1.5 + 1e-6
        else:
            assert_close(f"step {t} position targets", task.targets_loco, ref)
            assert 3.0 < float(ref.abs().max()) <= 3.1415928
        st = scene_cpu.generate(t, p.goal_quaternions)
        scene_cpu.load(st); scene_gpu.load(st)
        p.post_physics_step(); task.post_physics_step()
        assert_close(f"step {t} obs", task.obs_buf, p.obs_buf)
        assert torch.equal(task.reset_buf.cpu(), p.reset_buf)


@pytest.mark.parametrize("n", [258, 8192, 262144])
def test_deferred_finish_matches_inline(n):
    """lm_task_set_defer_finish: the finish pass of post(t) as an extra CTA of pre(t+1) (<= 1,024 tiles) or the stand-alone kernel behind
    it (2,048 tiles), the reset ranks taken from the per-tile counts — against the default mode (last-CTA tail inside the post kernel)
    on the same inputs: every buffer bit-identical after each pre step, extras / counters / num_reset bit-identical once the pass ran
    (also when lm_task_post_finish runs it early and the next pre launch carries it again)."""
    from locomanipulationrl_b200.tasks.base import LocoManipulationTask
    cfg = configs.get_task_config("JointLocomanipulationVertical")
    tasks, scenes = [], []
    for defer in (False, True):
        sc = synthetic.SyntheticScene(cfg, n, device="cuda:0", seed=3)
        t = LocoManipulationTask(cfg, n, scene=sc, device="cuda:0", seed=3)
        if defer:
            t.set_defer_finish(True)
        tasks.append(t); scenes.append(sc)
    acts, unis = rollout_inputs(n, 9, seed=21)
    ref, dfr = tasks
    fields = ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "successes", "consecutive_successes", "goal_reset_buf", "goal_quaternions")
    for t in range(9):
        st = scenes[0].generate(t)
        for task, sc in zip(tasks, scenes):
            task.pre_physics_step(acts[t].cuda(), unis[t].cuda())
        torch.cuda.synchronize()
        # the pass of post(t-1) has run inside (or behind) this pre launch
        if t > 0:
            assert torch.equal(dfr._extras_buf, ref._extras_buf) and torch.equal(dfr._counters, ref._counters), f"step {t}: extras / counters"
            assert torch.equal(dfr.num_reset_buf, ref.num_reset_buf), f"step {t}: num_reset"
            k = int(ref.num_reset_buf[0])
            assert torch.equal(dfr.reset_env_ids()[:k], ref.reset_env_ids()[:k]), f"step {t}: reset id list"
        for f in fields:
            assert torch.equal(getattr(dfr, f), getattr(ref, f)), f"step {t}: {f} after pre"
        for grp in ("loco", "mani"):
            for k_, v in getattr(scenes[0], grp).items():
                assert torch.equal(getattr(scenes[1], grp)[k_], v), f"step {t}: scene.{grp}.{k_} after pre"
        for task, sc in zip(tasks, scenes):
            sc.load(st)
            task.post_physics_step()
        if t % 3 == 2:                      # an early finish (rollout boundary): the next pre launch carries the pass again and must not act twice
            dfr.post_finish()
            torch.cuda.synchronize()
            assert torch.equal(dfr._extras_buf, ref._extras_buf) and torch.equal(dfr._counters, ref._counters), f"step {t}: early finish"
        for f in ("obs_buf", "rew_buf", "reset_buf", "progress_buf"):
            assert torch.equal(getattr(dfr, f), getattr(ref, f)), f"step {t}: {f} after post"


@pytest.mark.parametrize("spg", [6, 48])
def test_bench_multi_step_graph_equals_per_step_graphs(spg):
    """bench.py replays ONE graph holding a whole rollout horizon (48 steps = 8 rotations of the six state sets; the PDL chain then
    spans the step boundaries); the per-step graphs are what the port tracks above.  Same seeds: every task / policy buffer
    bit-identical at the graph boundaries."""
    import bench
    dev = torch.device("cuda:0")
    S = bench.STATE_SETS
    pipes = []
    for k in (1, spg):
        pipe = bench.Pipeline("cotrain64k", dev, rank=0)
        pipe.capture(steps_per_graph=k)
        pipes.append(pipe)
    a, b = pipes
    assert len(a.graphs) == S and len(b.graphs) == 1
    for rot in range(3 if spg == S else 2):
        for t in range(rot * spg, (rot + 1) * spg):
            a.step(t); b.step(t)
        for p_ in pipes:
            p_.task.post_finish()
        torch.cuda.synchronize()
        for f in ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "successes", "consecutive_successes", "goal_reset_buf", "goal_quaternions",
                  "_extras_buf", "_counters", "_actions_ring", "num_reset_buf", "last_base_tip_positions", "last_joint_velocities",
                  "targets_loco", "targets_mani"):
            assert torch.equal(getattr(a.task, f), getattr(b.task, f)), f"rotation {rot}: {f}"
        assert torch.equal(a.mu, b.mu) and torch.equal(a.value, b.value), f"rotation {rot}: policy outputs"
        for i in range(S):
            for grp in ("loco", "mani"):
                for k, v in getattr(a.scenes[i], grp).items():
                    assert torch.equal(getattr(b.scenes[i], grp)[k], v), f"rotation {rot}: scene[{i}].{grp}.{k}"


@pytest.mark.parametrize("n", [258, 4096, 20000, 80000])
def test_early_reset_equals_late_reset(n):
    """lm_task_set_early_reset (+ deferred finish): the reset half of pre_physics_step as k_pre_early — at most one CTA per SM, up to
    four tiles per pass (80,000 envs = 626 tiles: two passes; 20,000: 158 tiles on 148 CTAs; 258: ragged two-tile halves) — released
    beside the policy kernel, against the same loop with the reset half behind it.  Every task buffer and the scene rows bit-identical
    after every step."""
    from locomanipulationrl_b200.policy.mlp import MlpActorCritic
    from locomanipulationrl_b200.tasks.base import LocoManipulationTask
    cfg = configs.get_task_config("JointLocomanipulationVertical")
    torch.manual_seed(5)
    policy = MlpActorCritic(device="cuda:0")
    with torch.no_grad():
        policy.mean_layer.weight.mul_(20.0)
    tasks, scenes, outs = [], [], []
    for early in (False, True):
        sc = synthetic.SyntheticScene(cfg, n, device="cuda:0", seed=8)
        t = LocoManipulationTask(cfg, n, scene=sc, device="cuda:0", seed=8)
        if early:
            t.set_defer_finish(True)
            t.set_early_reset(True)
        tasks.append(t); scenes.append(sc)
        outs.append((torch.zeros(n, 12, device="cuda"), torch.zeros(n, 1, device="cuda")))
    _, unis = rollout_inputs(n, 10, seed=17)
    for task in tasks:
        task.pre_physics_step(torch.zeros(n, 12, device="cuda"), unis[0].cuda())
    fields = ("obs_buf", "rew_buf", "reset_buf", "progress_buf", "successes", "consecutive_successes", "goal_reset_buf", "goal_quaternions",
              "_actions_ring", "last_base_tip_positions", "last_joint_velocities", "targets_loco", "targets_mani", "num_reset_buf")
    with torch.no_grad():
        for t in range(9):
            st = scenes[0].generate(t)
            for task, sc, (mu, v) in zip(tasks, scenes, outs):
                sc.load(st)
                task.post_physics_step()
                policy(task.obs_buf, out=(mu, v), sink=task.action_sink())
                task.pre_physics_step(None, unis[t + 1].cuda())
            tasks[1].post_finish()
            torch.cuda.synchronize()
            for f in fields + ("_extras_buf", "_counters"):
                assert torch.equal(getattr(tasks[0], f), getattr(tasks[1], f)), f"step {t}: {f}"
            k = int(tasks[0].num_reset_buf[0])
            assert torch.equal(tasks[0].reset_env_ids()[:k], tasks[1].reset_env_ids()[:k]), f"step {t}: reset id list"
            assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), f"step {t}: policy outputs"
            for grp in ("loco", "mani"):
                for k_, v_ in getattr(scenes[0], grp).items():
                    assert torch.equal(getattr(scenes[1], grp)[k_], v_), f"step {t}: scene.{grp}.{k_}"
