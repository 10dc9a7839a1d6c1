"""CPU tests: the oracle port against the golden vectors generated from the UNMODIFIED reference
(oracle/make_golden.py), so the checker the GPU tests rely on is itself pinned."""
import os

import numpy as np
import pytest
import torch

from locomanipulationrl_b200 import configs, synthetic
from oracle import isaac_torch as it
from oracle import port, port_policy
from tests.common import GOLDEN_DIR, GOLDEN_FULL_STEPS, GOLDEN_N, GOLDEN_SEED, GOLDEN_STEPS, golden_steps, rollout_inputs


def _load(name):
    return np.load(os.path.join(GOLDEN_DIR, name))


@pytest.mark.parametrize("name", sorted(configs.TASK_CONFIGS))
def test_task_port_matches_reference_bit_exact(name):
    g = _load(f"task_{name}.npz")
    cfg = configs.get_task_config(name)
    scene = synthetic.SyntheticScene(cfg, GOLDEN_N, seed=GOLDEN_SEED)
    task = port.TaskPort(cfg, scene)
    steps, full_steps = golden_steps(name)
    acts, unis = rollout_inputs(GOLDEN_N, steps)
    saw_reset_after_first, saw_goal_reset = False, False
    pd = cfg.control_mode == configs.CONTROL_PD
    for t in range(steps):
        task.pre_physics_step(acts[t].clone(), unis[t], (lambda i, t=t: scene.substep(t, i)) if pd else None)
        assert np.array_equal(task.reset_env_ids.numpy(), g[f"s{t}_reset_ids"]), f"step {t} reset ids"
        assert np.array_equal(task.goal_quaternions.numpy(), g[f"s{t}_goal"]), f"step {t} goals"
        if pd:
            assert np.array_equal(task.torque.numpy(), g[f"s{t}_torque"]), f"step {t} mean torque"
        if t in full_steps:
            if pd:
                assert np.array_equal(task.current_joint_position_targets.numpy(), g[f"s{t}_pd_targets"])
                assert np.array_equal(task.current_joint_position_targets_se.numpy(), g[f"s{t}_pd_targets_se"])
            for grp in ("loco", "mani"):
                st = getattr(scene, grp)
                if st:
                    h = scene.num_scenes
                    rows = slice(h, 2 * h) if (grp == "mani" and cfg.kind == configs.KIND_COTRAIN) else slice(0, h)
                    tg = task.applied_torque[rows] if pd else task.targets[grp]
                    assert np.array_equal(tg.numpy(), g[f"s{t}_{grp}_targets"])
                    for k in ("dof_pos", "dof_vel", "indicator_pos", "indicator_quat"):
                        assert np.array_equal(st[k].numpy(), g[f"s{t}_{grp}_{k}"]), f"step {t} {grp}.{k}"
        scene.step(t, task.goal_quaternions)
        obs, rew, reset, extras = task.post_physics_step()
        if t in full_steps:
            assert np.array_equal(obs.numpy(), g[f"s{t}_obs"]), f"step {t} obs"
            if cfg.num_states == 93:
                assert np.array_equal(task.states_buf.numpy(), g[f"s{t}_states"])
            assert np.array_equal(task.last_base_tip_positions.numpy(), g[f"s{t}_last_base_tip_positions"])
        assert np.array_equal(rew.numpy(), g[f"s{t}_rew"]), f"step {t} rew"
        assert np.array_equal(reset.numpy(), g[f"s{t}_reset"].astype(np.int64)), f"step {t} reset"
        for k in ("progress_buf", "successes", "consecutive_successes", "goal_reset_buf"):
            assert np.array_equal(getattr(task, k).numpy(), g[f"s{t}_{k}"].astype(np.int64)), f"step {t} {k}"
        for k, v in extras.items():
            assert np.float32(float(v)) == g[f"s{t}_extras/{k}"], f"step {t} extras {k}"
        saw_reset_after_first |= t > 0 and len(g[f"s{t}_reset_ids"]) > 0
        saw_goal_reset |= bool(g[f"s{t}_goal_reset_buf"].any())
    assert saw_reset_after_first and saw_goal_reset, "fixture must exercise resets and goal resets"


def test_math_port_matches_reference():
    g = _load("math_reference.npz")
    q, q2, t, v = (torch.from_numpy(g[k]) for k in ("q", "q2", "t", "v"))
    assert np.array_equal(port.transform_vectors(q, t, v).numpy(), g["transform"])
    assert np.array_equal(port.inverse_transform_vectors(q, t, v).numpy(), g["inverse_transform"])
    assert np.array_equal(port.rotate_orientations(q, q2).numpy(), g["rotate"])
    assert np.array_equal(port.inverse_rotate_orientations(q, q2).numpy(), g["inverse_rotate"])
    # the reference's only executable check (utils/math.py:214-218)
    ex = port.inverse_rotate_orientations(torch.tensor([-0.5, -0.5, 0.5, 0.5]).repeat(2, 1), torch.tensor([0.7071, 0, 0, 0.7071]).repeat(2, 1))
    assert np.array_equal(ex.numpy(), g["main_example"])


def test_gnn_port_matches_reference():
    g = _load("gnn_reference.npz")
    p = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("p/")}
    mu, value = port_policy.gnn_forward(p, torch.from_numpy(g["obs"]))
    np.testing.assert_allclose(mu.numpy(), g["mu"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(value.numpy(), g["value"], rtol=1e-6, atol=1e-6)
    assert port_policy.limb_edges().shape == (2, 24)


def test_isaac_restatement_properties():
    """The Isaac helpers are third-party and unpinned: check algebraic identities and agreement with SciPy."""
    from scipy.spatial.transform import Rotation
    torch.manual_seed(0)
    q = torch.randn(256, 4, dtype=torch.float64)
    q = (q / q.norm(dim=-1, keepdim=True)).float()
    r = torch.randn(256, 4, dtype=torch.float64)
    r = (r / r.norm(dim=-1, keepdim=True)).float()
    v = torch.randn(256, 3)
    ident = it.quat_mul(q, it.quat_conjugate(q))
    assert torch.allclose(ident, torch.tensor([1.0, 0, 0, 0]).expand_as(ident), atol=1e-6)
    assert torch.allclose(it.quat_mul(q, r).norm(dim=-1), torch.ones(256), atol=1e-6)
    assert torch.allclose(it.quat_rotate_inverse(q, it.quat_rotate(q, v)), v, atol=1e-5)
    R = Rotation.from_quat(q[:, [1, 2, 3, 0]].numpy())
    assert np.allclose(it.quat_rotate(q, v).numpy(), R.apply(v.numpy()), atol=1e-5)
    assert np.allclose(it.quat_axis(q, 2).numpy(), R.as_matrix()[:, :, 2], atol=1e-6)
    qr = (R * Rotation.from_quat(r[:, [1, 2, 3, 0]].numpy())).as_quat()[:, [3, 0, 1, 2]]
    mine = it.quat_mul(q, r).numpy()
    sign = np.sign((qr * mine).sum(-1, keepdims=True))
    assert np.allclose(mine * sign, qr, atol=1e-6)
    e = torch.rand(64, 3) * 2 - 1
    qe = it.quat_from_euler_xyz(e[:, 0], e[:, 1], e[:, 2]).numpy()
    qs = Rotation.from_euler("xyz", e.numpy()).as_quat()[:, [3, 0, 1, 2]]
    assert np.allclose(qe * np.sign((qe * qs).sum(-1, keepdims=True)), qs, atol=1e-6)
    lo, hi = torch.full((12,), -3.0), torch.full((12,), 3.0)
    assert torch.equal(it.unscale_transform(torch.tensor([[0.5] * 12]), lo, hi), torch.full((1, 12), 1.5))


def test_gae_ports_closed_form():
    T, N, g, lam = 16, 8, 0.99, 0.95
    r = torch.ones(T, N)
    v = torch.zeros(T, N)
    d = torch.zeros(T, N)
    adv, ret = port_policy.gae_rlgames(r, v, d, torch.zeros(N), torch.zeros(N), g, lam)
    expect = torch.tensor([sum((g * lam) ** k for k in range(T - t)) for t in range(T)])
    assert torch.allclose(adv[:, 0], expect, rtol=1e-5)
    adv2, _ = port_policy.gae_skrl(r, v, d, torch.zeros(N), g, lam)
    assert torch.allclose(adv2[:, 0], expect, rtol=1e-5)
    # a done cuts the bootstrap: rl_games masks with the NEXT step's done, skrl with the current step's
    d[5] = 1.0
    a1, _ = port_policy.gae_rlgames(r, v, d, torch.zeros(N), torch.zeros(N), g, lam)
    a2, _ = port_policy.gae_skrl(r, v, d, torch.zeros(N), g, lam)
    assert torch.allclose(a1[4], torch.ones(N)) and torch.allclose(a2[5], torch.ones(N))
    assert torch.equal(ret, adv + v)


def test_synthetic_scene_shapes_and_rates():
    cfg = configs.get_task_config("JointLocomanipulationVertical")
    scene = synthetic.SyntheticScene(cfg, 512)
    scene.step(0)
    assert scene.env_pos.shape == (256, 3)
    assert scene.loco["dof_pos"].shape == (256, 20) and scene.loco["tip_pos"].shape == (1024, 3)
    assert scene.mani["knee_pos"].shape == (2048, 3) and scene.mani["plate_quat"].shape == (256, 4)
    task = port.TaskPort(cfg, scene)
    task.pre_physics_step(torch.zeros(512, 12), torch.rand(512, 3))
    assert int(task.reset_buf.sum()) == 0 and task.reset_env_ids.numel() == 512      # reset storm on the first step
    scene.step(0, task.goal_quaternions)
    task.post_physics_step()
    assert 0 < int(task.reset_buf.sum()) < 512 * 0.2
    with pytest.raises(AssertionError):
        synthetic.SyntheticScene(cfg, 511)     # "Number of envs must be a multiplier of 2." (JLV:135-136)
