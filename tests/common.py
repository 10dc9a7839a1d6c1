"""Shared drivers for the parity tests: seeded rollouts of the oracle port and of the CUDA task."""
from __future__ import annotations

import os

import numpy as np
import torch

from locomanipulationrl_b200 import configs, synthetic

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_N, GOLDEN_STEPS, GOLDEN_SEED = 50, 20, 7
GOLDEN_FULL_STEPS = (0, 1, 2, 17, 18, 19)
# fixtures of tasks whose goal reset needs more than 20 steps (custom controller: max_consecutive_successes = 20)
GOLDEN_STEPS_OVERRIDE = {"QuadrupedPoseControlCustomController": 28, "QuadrupedManipulatePlateCustomController": 28}


def golden_steps(name):
    """(number of steps, steps whose float tensors are stored) of the fixture of task ``name``."""
    n = GOLDEN_STEPS_OVERRIDE.get(name, GOLDEN_STEPS)
    return n, (0, 1, 2, n - 3, n - 2, n - 1)

# fp32 tolerance of BASELINE.json's north_star ("within 1e-5 relative"), read norm-wise with an absolute floor
# because several observation components cross zero (SURVEY.md §7 "hard parts").
RTOL, ATOL = 1e-5, 2e-6
# a mask may differ from the oracle only where the oracle's own margin to the threshold is below this (metres / radians);
# differences come from <= 1 ulp libm-level deviations (asin, sin, cos), see DESIGN.md "numerics".
MARGIN_EPS = 2e-6


def rollout_inputs(n, steps, seed=GOLDEN_SEED):
    g = torch.Generator().manual_seed(seed)
    acts = [(torch.rand(n, 12, generator=g) * 2 - 1) for _ in range(steps)]
    unis = [torch.rand(n, 3, generator=g) for _ in range(steps)]
    return acts, unis


def rot_reward_conditioning(p, ulps=4.0):
    """Forward-error allowance for the orientation reward 0.5 / (2 asin(min(x, 1)) + 0.1) of the reference (QPC:431-432):
    d asin / dx = 1 / sqrt(1 - x^2) is unbounded as x -> 1 (180-degree orientation error), so a 2-ulp perturbation of
    x = |quat_diff.xyz| (goal quaternions differ by <= 1 ulp between libm implementations; ``ulps`` = 4 by default, i.e. 2.4e-7)
    moves rot_dist by 2 * 2.4e-7 / sqrt(1 - x^2).  Away from x ~ 1 this term is < 1e-7 and irrelevant."""
    x = torch.clamp(p.inter["quat_diff_norm"].double(), max=1.0)
    d_rd = 2.0 * (ulps * 5.96e-8) / torch.sqrt(torch.clamp(1.0 - x * x, min=1.2e-7))      # ulp(x) = 2^-24 just below 1
    d_rew = 0.5 / (p.inter["rot_dist"].double().abs() + p.cfg.rot_eps) ** 2
    return d_rew * d_rd


def assert_close(name, got, ref, rtol=RTOL, atol=ATOL, rows=None, scale=None, extra_atol=None):
    """|got - ref| <= atol + rtol * max(|ref|, scale).  ``scale`` (per element) is the magnitude of the terms a value was
    summed from: the reward is rot_rew (up to 5) minus penalties and often cancels to ~0, so its fp32 error is relative
    to the terms, not to the total."""
    got = torch.as_tensor(got).detach().cpu().double()
    ref = torch.as_tensor(ref).detach().cpu().double()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    mag = ref.abs() if scale is None else torch.maximum(ref.abs(), torch.as_tensor(scale).detach().cpu().double())
    if rows is not None:
        got, ref, mag = got[rows], ref[rows], mag[rows]
    err = (got - ref).abs()
    tol = atol + rtol * mag
    if extra_atol is not None:
        ex = torch.as_tensor(extra_atol).detach().cpu().double()
        tol = tol + (ex[rows] if rows is not None else ex)
    bad = err > tol
    if bad.any():
        i = torch.nonzero(bad)[0].tolist()
        raise AssertionError(f"{name}: {int(bad.sum())} elements out of tolerance; first at {i}: got {got[tuple(i)].item():.9g} "
                             f"ref {ref[tuple(i)].item():.9g} (max err {err.max().item():.3g})")
    return float(err.max()) if err.numel() else 0.0


class CudaRollout:
    """Drives the CUDA task and the oracle port side by side on identical synthetic state."""

    def __init__(self, name, n, seed=GOLDEN_SEED, device="cuda:0", knobs=None, cfg=None):
        from locomanipulationrl_b200.tasks.base import LocoManipulationTask
        from oracle import port
        self.cfg = cfg or configs.get_task_config(name)
        self.n = n
        self.scene_cpu = synthetic.SyntheticScene(self.cfg, n, seed=seed, knobs=knobs)
        self.scene_gpu = synthetic.SyntheticScene(self.cfg, n, device=device, seed=seed, knobs=knobs)
        self.port = port.TaskPort(self.cfg, self.scene_cpu)
        self.task = LocoManipulationTask(self.cfg, n, scene=self.scene_gpu, device=device)
        self.exempt_events = 0
        self.max_err = {}

    def _track(self, k, e):
        self.max_err[k] = max(self.max_err.get(k, 0.0), e)

    def step(self, t, actions, uniforms, check=True):
        p, task = self.port, self.task
        dev = task._device
        pd = self.cfg.control_mode == configs.CONTROL_PD
        if pd:   # the explicit-PD files step the simulation inside pre_physics_step: same seeded sub-step joint states on both sides
            sub = [self.scene_cpu.substep_joint_state(t, i) for i in range(self.cfg.pd_control_decimal)]
            p.pre_physics_step(actions.clone(), uniforms, lambda i: self.scene_cpu.load(sub[i]))
            task.pre_physics_step(actions.to(dev), uniforms.to(dev), lambda i: self.scene_gpu.load(sub[i]))
        else:
            p.pre_physics_step(actions.clone(), uniforms)
            task.pre_physics_step(actions.to(dev), uniforms.to(dev))
        if check and pd:
            self._track("torque", assert_close(f"step {t} mean torque", task.torque, p.torque))
            self._track("torque", assert_close(f"step {t} applied torque", task.applied_torque, p.applied_torque))
            self._track("pd_targets", assert_close(f"step {t} pd targets", task.current_joint_position_targets, p.current_joint_position_targets))
            self._track("pd_targets", assert_close(f"step {t} pd targets_se", task.current_joint_position_targets_se, p.current_joint_position_targets_se))
            self._track("pd_state", assert_close(f"step {t} pd joint_positions", task.joint_positions, p.joint_positions))
            self._track("pd_state", assert_close(f"step {t} pd joint_velocities", task.joint_velocities, p.joint_velocities))
        if check:
            ids_ref = p.reset_env_ids
            ids = task.reset_env_ids().cpu()
            assert torch.equal(ids, ids_ref), f"step {t}: reset env-id list differs ({ids.numel()} vs {ids_ref.numel()})"
            nr = task.num_reset_buf.cpu()
            h = p.half if self.cfg.kind == configs.KIND_COTRAIN else (self.n if self.cfg.kind == configs.KIND_LOCO else 0)
            assert int(nr[0]) == ids_ref.numel() and int(nr[1]) == int((ids_ref < h).sum()) if self.cfg.kind != configs.KIND_MANI else True
            self._track("goal", assert_close(f"step {t} goal_quaternions", task.goal_quaternions, p.goal_quaternions))
            for grp in ("loco", "mani"):
                ref_state = getattr(self.scene_cpu, grp)
                if not ref_state:
                    continue
                got_state = getattr(self.scene_gpu, grp)
                for k, v in ref_state.items():
                    self._track("scene", assert_close(f"step {t} scene.{grp}.{k}", got_state[k], v))
                if not pd:
                    tg = task.targets_loco if grp == "loco" else task.targets_mani
                    self._track("targets", assert_close(f"step {t} targets_{grp}", tg, p.targets[grp]))
            assert torch.equal(task.reset_buf.cpu(), p.reset_buf), f"step {t}: reset_buf after pre differs"
            assert torch.equal(task.progress_buf.cpu(), p.progress_buf)
        # identical "physics" on both sides, generated from the ORACLE's goals
        states = self.scene_cpu.generate(t, p.goal_quaternions)
        self.scene_cpu.load(states)
        self.scene_gpu.load(states)
        p.post_physics_step()
        task.post_physics_step()
        if check:
            self.compare_post(t)

    def compare_post(self, t):
        p, task = self.port, self.task
        margins = p.threshold_margins()
        int_pairs = {"reset_buf": (task.reset_buf, p.reset_buf), "successes": (task.successes, p.successes),
                     "consecutive_successes": (task.consecutive_successes, p.consecutive_successes),
                     "goal_reset_buf": (task.goal_reset_buf, p.goal_reset_buf), "progress_buf": (task.progress_buf, p.progress_buf)}
        bad = torch.zeros(self.n, dtype=torch.bool)
        for k, (got, ref) in int_pairs.items():
            got = got.cpu()
            assert got.dtype == ref.dtype == torch.long, f"{k} dtype"
            bad |= got != ref
        if bad.any():
            rows = torch.nonzero(bad).squeeze(-1)
            far = margins[rows] > MARGIN_EPS
            assert not far.any(), (f"step {t}: mask mismatch away from any threshold at envs {rows[far].tolist()} "
                                   f"(margins {margins[rows][far].tolist()})")
            self.exempt_events += int(rows.numel())
            for k, (got, ref) in int_pairs.items():   # re-synchronise the diverged envs to the oracle
                got[rows.to(got.device)] = ref[rows].to(got.device)
        ok = ~bad
        self._track("obs", assert_close(f"step {t} obs_buf", task.obs_buf, p.obs_buf))
        if self.cfg.num_states == 93:
            self._track("states", assert_close(f"step {t} states_buf", task.states_buf, p.states_buf))
        self._track("rew", assert_close(f"step {t} rew_buf", task.rew_buf, p.rew_buf, rows=ok, scale=p.inter["rew_scale"],
                                        extra_atol=rot_reward_conditioning(p, getattr(self, "cond_ulps", 4.0))))
        self._track("last_tip", assert_close(f"step {t} last_base_tip", task.last_base_tip_positions, p.last_base_tip_positions))
        self._track("last_jv", assert_close(f"step {t} last_joint_vel", task.last_joint_velocities,
                                            torch.cat([p.last_joint_velocities[g] for g in ("loco", "mani") if getattr(self.scene_cpu, g)], dim=0)))
        self._track("last_actions", assert_close(f"step {t} last_actions", task.last_actions, p.last_actions, atol=0, rtol=0))
        if self.cfg.control_mode == configs.CONTROL_PD:
            self._track("pd_state", assert_close(f"step {t} pd joint_positions (post)", task.joint_positions, p.joint_positions))
            self._track("pd_state", assert_close(f"step {t} pd last_targets", task.last_joint_position_targets, p.last_joint_position_targets))
        if self.cfg.custom_controller and self.cfg.cc_rot_dist_term:
            self._track("last_rot_dist", assert_close(f"step {t} last_rot_dist", task.last_rot_dist, p.last_rot_dist, rtol=1e-6, atol=1e-6))
        if self.exempt_events == 0:
            for k, v in p.extras.items():
                self._track("extras", assert_close(f"step {t} extras[{k}]", task.extras[k], v, rtol=2e-5, atol=1e-6))
            c = task._counters.cpu()
            names = ["num_successes", "num_resets", "num_successes_loco", "num_resets_loco", "num_successes_mani", "num_resets_mani"]
            for i, k in enumerate(names[:6 if self.cfg.kind == configs.KIND_COTRAIN else 2]):
                assert int(c[i]) == int(p.counters[k]), f"step {t}: counter {k}: {int(c[i])} vs {int(p.counters[k])}"
