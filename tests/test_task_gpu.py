"""GPU parity tests of the task kernels, through the C ABI, against the oracle port and the golden vectors."""
import os

import numpy as np
import pytest
import torch

from locomanipulationrl_b200 import configs, synthetic
from tests.common import (ATOL, GOLDEN_DIR, GOLDEN_FULL_STEPS, GOLDEN_N, GOLDEN_SEED, GOLDEN_STEPS, RTOL, CudaRollout, golden_steps,
                          assert_close, rollout_inputs)

pytestmark = pytest.mark.gpu
ALL_TASKS = sorted(configs.TASK_CONFIGS)


@pytest.mark.parametrize("name", ALL_TASKS)
def test_cuda_matches_golden_reference_vectors(name):
    """CUDA vs the vectors produced by the unmodified reference (N = 50: ragged, non-TMA tile path)."""
    from locomanipulationrl_b200.tasks.base import LocoManipulationTask
    g = np.load(os.path.join(GOLDEN_DIR, f"task_{name}.npz"))
    cfg = configs.get_task_config(name)
    scene = synthetic.SyntheticScene(cfg, GOLDEN_N, device="cuda:0", seed=GOLDEN_SEED)
    task = LocoManipulationTask(cfg, GOLDEN_N, scene=scene)
    steps, full_steps = golden_steps(name)
    acts, unis = rollout_inputs(GOLDEN_N, steps)
    pd = cfg.control_mode == configs.CONTROL_PD
    for t in range(steps):
        task.pre_physics_step(acts[t].cuda(), unis[t].cuda(), (lambda i, t=t: scene.substep(t, i)) if pd else None)
        assert np.array_equal(task.reset_env_ids().cpu().numpy(), g[f"s{t}_reset_ids"]), f"step {t}: reset ids"
        assert_close(f"step {t} goal", task.goal_quaternions, g[f"s{t}_goal"])
        if pd:
            assert_close(f"step {t} mean torque", task.torque, g[f"s{t}_torque"])
        if t in full_steps:
            if pd:
                assert_close(f"step {t} pd targets", task.current_joint_position_targets, g[f"s{t}_pd_targets"])
                assert_close(f"step {t} pd targets_se", task.current_joint_position_targets_se, g[f"s{t}_pd_targets_se"])
            h = scene.num_scenes
            for grp, tg in (("loco", task.targets_loco), ("mani", task.targets_mani)):
                st = getattr(scene, grp)
                if st:
                    if pd:
                        tg = task.applied_torque[h:2 * h] if (grp == "mani" and cfg.kind == configs.KIND_COTRAIN) else task.applied_torque[:h]
                    assert_close(f"step {t} {grp} targets", tg, g[f"s{t}_{grp}_targets"])
                    for k in ("dof_pos", "dof_vel", "indicator_pos", "indicator_quat"):
                        assert_close(f"step {t} {grp}.{k}", st[k], g[f"s{t}_{grp}_{k}"])
        scene.step(t, torch.from_numpy(g[f"s{t}_goal"]))
        obs, rew, reset, extras = task.post_physics_step()
        if t in full_steps:
            assert_close(f"step {t} obs", obs, g[f"s{t}_obs"])
            if cfg.num_states == 93:
                assert_close(f"step {t} states", task.get_states(), g[f"s{t}_states"])
            assert_close(f"step {t} last tips", task.last_base_tip_positions, g[f"s{t}_last_base_tip_positions"])
        assert_close(f"step {t} rew", rew, g[f"s{t}_rew"], scale=torch.full((GOLDEN_N,), 5.0))   # rot_rew term is up to 0.5/0.1
        assert np.array_equal(reset.cpu().numpy(), g[f"s{t}_reset"].astype(np.int64)), f"step {t}: reset mask"
        for k in ("progress_buf", "successes", "consecutive_successes", "goal_reset_buf"):
            assert np.array_equal(getattr(task, k).cpu().numpy(), g[f"s{t}_{k}"].astype(np.int64)), f"step {t}: {k}"
        for k in extras:
            assert_close(f"step {t} extras {k}", extras[k], g[f"s{t}_extras/{k}"], rtol=2e-5, atol=1e-6)
    assert obs.dtype == torch.float32 and reset.dtype == torch.long and task.progress_buf.dtype == torch.long


@pytest.mark.parametrize("name,n", [("QuadrupedPoseControl", 4096), ("QuadrupedManipulatePlate", 16384),
                                    ("JointLocomanipulationVertical", 8192), ("JointLocomanipulation", 4096),
                                    ("QuadrupedPoseControlVertical", 1024), ("QuadrupedManipulatePlateVertical", 1024),
                                    ("QuadrupedPoseControlPositionControl", 2048), ("QuadrupedManipulatePlatePositionControl", 2048),
                                    ("JointLocomanipulationPositionControl", 4096), ("QuadrupedPoseControlCustomController", 4096),
                                    ("QuadrupedManipulatePlateCustomController", 2048)])
def test_cuda_matches_oracle_port_rollout(name, n):
    """BASELINE.json configs 1-3 shapes (config 3 at 8192 to keep the CPU oracle quick): 24-step rollouts, TMA tile path."""
    ro = CudaRollout(name, n)
    acts, unis = rollout_inputs(n, 24, seed=11)
    for t in range(24):
        ro.step(t, acts[t], unis[t])
    assert ro.exempt_events <= 2, f"{ro.exempt_events} near-threshold mask exemptions"
    print(name, n, "max abs err:", {k: f"{v:.2e}" for k, v in ro.max_err.items()}, "exempt:", ro.exempt_events)


def test_custom_controller_optional_terms():
    """quadruped_pose_control_custom_controller.py ships rot_dist_decreasing_reward_scale = 0 (``# 50``) and clip_reward = False:
    switch both on (the port is pinned to the reference class with these values too) and track the port."""
    cfg = configs.get_task_config("QuadrupedPoseControlCustomController").with_(rot_dist_decreasing_reward_scale=50.0, clip_reward=True)
    ro = CudaRollout(cfg.name, 1000, cfg=cfg)
    acts, unis = rollout_inputs(1000, 30, seed=5)
    for t in range(30):
        ro.step(t, acts[t], unis[t])
    assert ro.exempt_events <= 2
    assert float(ro.port.rew_buf.min()) >= 0.0 and float(ro.task.rew_buf.min()) >= 0.0
    assert abs(float(ro.port.extras["env/rewards/rot_dist_decreasing_reward"])) > 1e-3        # the term was live


@pytest.mark.parametrize("n", [2, 6, 130, 258])
def test_ragged_sizes_cotrain(n):
    ro = CudaRollout("JointLocomanipulationVertical", n)
    acts, unis = rollout_inputs(n, 6, seed=3)
    for t in range(6):
        ro.step(t, acts[t], unis[t])


def test_edge_cases_termination_and_timeout():
    """Force each termination cause (QPC:562-616) and the episode timeout, and a reset storm via task.reset()."""
    n = 256
    ro = CudaRollout("QuadrupedPoseControl", n, knobs=synthetic.SynthKnobs(p_low_knee=0.2, p_fallen=0.2, base_rp_range=0.6))
    acts, unis = rollout_inputs(n, 8, seed=5)
    for t in range(3):
        ro.step(t, acts[t], unis[t])
    ro.port.progress_buf[:] = 297
    ro.task.progress_buf.fill_(297)
    ro.step(3, acts[3], unis[3])
    ro.step(4, acts[4], unis[4])                     # progress hits max_episode_length - 1 = 299 for survivors
    survivors = ro.task.progress_buf == 299
    assert int(survivors.sum()) > 0 and bool((ro.task.reset_buf[survivors] == 1).all())
    ro.port.reset_buf[:] = 1                         # then flag everyone
    ro.task.reset_buf.fill_(1)
    ro.step(5, acts[5], unis[5])                     # everyone resets (storm); id lists are compared inside step()
    assert ro.port.reset_env_ids.numel() == n
    ro.port.reset_buf = torch.ones_like(ro.port.reset_buf)   # rl_task.py:222-225
    ro.task.reset()
    ro.step(6, acts[6], unis[6])
    assert ro.port.reset_env_ids.numel() == n
    ro.step(7, acts[7], unis[7])


def test_unclamped_actions_are_clamped_like_the_adapter():
    """VecEnvRLGames.step clamps to ±clipActions before the task sees the actions (envs/vec_env_rlgames.py:60)."""
    n = 512
    ro = CudaRollout("QuadrupedManipulatePlate", n)
    g = torch.Generator().manual_seed(2)
    for t in range(3):
        raw = torch.randn(n, 12, generator=g) * 2.0
        u = torch.rand(n, 3, generator=g)
        clamped = torch.clamp(raw, -1.0, 1.0)
        ro.port.pre_physics_step(clamped.clone(), u)
        ro.task.pre_physics_step(raw.cuda(), u.cuda())
        assert torch.equal(ro.task.current_actions.cpu(), clamped)
        st = ro.scene_cpu.generate(t, ro.port.goal_quaternions)
        ro.scene_cpu.load(st); ro.scene_gpu.load(st)
        ro.port.post_physics_step(); ro.task.post_physics_step()
        ro.compare_post(t)


def test_fused_observation_clamp_and_clamp_copy():
    from locomanipulationrl_b200 import _lib
    from locomanipulationrl_b200.tasks.base import LocoManipulationTask
    n = 1024
    cfg = configs.get_task_config("QuadrupedPoseControl")
    s1 = synthetic.SyntheticScene(cfg, n, device="cuda:0")
    s2 = synthetic.SyntheticScene(cfg, n, device="cuda:0")
    a = LocoManipulationTask(cfg, n, scene=s1)
    b = LocoManipulationTask(cfg, n, scene=s2, fuse_obs_clamp=True)
    acts, unis = rollout_inputs(n, 2, seed=9)
    for t in range(2):
        for task, sc in ((a, s1), (b, s2)):
            task.pre_physics_step(acts[t].cuda() * 3, unis[t].cuda())
            sc.step(t, task.goal_quaternions)
            task.post_physics_step()
    assert torch.equal(torch.clamp(a.obs_buf, -5.0, 5.0), b.obs_buf)
    assert torch.equal(torch.clamp(a.states_buf, -5.0, 5.0), b.states_buf)
    out = torch.empty_like(a.obs_buf)
    lib = _lib.load()
    _lib.check(lib.lm_clamp_copy(a.obs_buf.data_ptr(), out.data_ptr(), a.obs_buf.numel(), 0.25, torch.cuda.current_stream().cuda_stream))
    assert torch.equal(out, torch.clamp(a.obs_buf, -0.25, 0.25))


def test_reset_idx_direct_call():
    n = 300
    ro = CudaRollout("JointLocomanipulationVertical", n)
    acts, unis = rollout_inputs(n, 2, seed=4)
    ro.step(0, acts[0], unis[0])
    ids = torch.tensor([3, 17, 149, 150, 151, 299])
    u = torch.rand(6, 3)
    ro.port.reset_idx(ids, u)
    ro.task.reset_idx(ids.cuda(), u.cuda())
    assert_close("goal after reset_idx", ro.task.goal_quaternions, ro.port.goal_quaternions)
    for grp in ("loco", "mani"):
        for k, v in getattr(ro.scene_cpu, grp).items():
            assert_close(f"{grp}.{k}", getattr(ro.scene_gpu, grp)[k], v)
    assert torch.equal(ro.task.reset_buf.cpu(), ro.port.reset_buf)
    ro.step(1, acts[1], unis[1])


def test_bad_arguments_return_errors():
    from locomanipulationrl_b200 import _lib
    from locomanipulationrl_b200.tasks.base import LocoManipulationTask
    cfg = configs.get_task_config("QuadrupedPoseControl")
    with pytest.raises(_lib.LmError):
        LocoManipulationTask(configs.get_task_config("JointLocomanipulation"), 7, device="cuda:0")   # odd env count
    n = 64
    scene = synthetic.SyntheticScene(cfg, n, device="cuda:0")
    task = LocoManipulationTask(cfg, n, scene=scene)
    misaligned = torch.zeros(n * 12 + 1, device="cuda")[1:].view(n, 12)
    with pytest.raises(_lib.LmError):
        task.pre_physics_step(misaligned, torch.rand(n, 3, device="cuda"))
    unbound = LocoManipulationTask(cfg, n)
    with pytest.raises(_lib.LmError):
        unbound.pre_physics_step(torch.zeros(n, 12, device="cuda"))
