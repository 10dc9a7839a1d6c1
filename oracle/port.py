"""TEST INFRASTRUCTURE — torch-CPU restatement of the reference's task hot path (the oracle "port").

One parameterised class restates what the reference spreads over six near-identical task files:

  QPC  quadruped_pose_control.py            pre :200-228  reset :230-299  obs :301-426  metrics :428-560  done :562-633
  QMP  quadruped_manipulate_plate.py        pre :213-236  reset :238-309  obs :311-435  metrics :437-567  done :569-652
  JLV  joint_locomanipulation_vertical.py   pre :248-270  reset :272-377  obs :379-521  metrics :523-653  done :655-802
  (+ quadruped_pose_control_vertical.py, quadruped_manipulate_plate_vertical.py, joint_locomanipulation.py: constant deltas)
  *_PC quadruped_pose_control_position_control.py:251-322,377-386,438-452 and its manipulation / co-train siblings
       (joint_locomanipulation_position_control.py:341-408): explicit PD loop, target rows in the observation
  robot/base/robot.py:276-321 (state read-back), :444-461 (take_action); objects/base/rigid_object.py:36-58;
  utils/math.py:7-193 (SciPy float64 quaternion<->matrix round trips + float32 bmm);
  tasks/base/rl_task.py:104-113,240-260 (buffers, post_physics_step driver).

It issues the SAME library calls as the reference (scipy Rotation in float64, torch.bmm / torch.norm /
torch.asin in float32, eager op-by-op arithmetic), so on one CPU its results equal the reference's bit
for bit; tests/test_oracle_vs_reference.py and the committed vectors in tests/golden/ pin that.
Because it mirrors the reference's CPU cost structure it is also what ``bench.py`` times as the
``cpu_baseline`` (kind "port") on the GPU box, where /root/reference does not exist.

The product never imports this file.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
from scipy.spatial.transform import Rotation

from locomanipulationrl_b200.configs import (CONTROL_EFFORT, CONTROL_PD, CONTROL_POSITION, CONTROL_VELOCITY, DEFAULT_BASE_TIP_POSITIONS,
                                             KIND_COTRAIN, KIND_LOCO, KIND_MANI, TaskConfig)

from . import isaac_torch as it


# ----------------------------------------------------------------------------- utils/math.py restated
def rot_from_quat(q: torch.Tensor) -> torch.Tensor:
    """math.py:7-19 — wxyz -> SciPy (xyzw) in float64 -> 3x3, cast to float32."""
    m = Rotation.from_quat(q[:, [1, 2, 3, 0]].detach().cpu().numpy()).as_matrix()
    out = torch.zeros(q.shape[0], 3, 3, dtype=torch.float32)
    out[:, :3, :3] = torch.from_numpy(m).float()
    return out


def quat_from_rot(m: torch.Tensor) -> torch.Tensor:
    """math.py:21-31."""
    q = Rotation.from_matrix(m.detach().cpu().numpy()).as_quat()[:, [3, 0, 1, 2]]
    return torch.from_numpy(q).float()


def _positive_w(q: torch.Tensor) -> torch.Tensor:
    return torch.where((q[:, 0] < 0.0).view(-1, 1), -q, q)


def rotate_orientations(rot: torch.Tensor, orient: torch.Tensor) -> torch.Tensor:
    """math.py:33-57."""
    return _positive_w(quat_from_rot(torch.bmm(rot_from_quat(rot), rot_from_quat(orient))))


def inverse_rotate_orientations(rot: torch.Tensor, orient: torch.Tensor) -> torch.Tensor:
    """math.py:59-83."""
    inv = torch.transpose(rot_from_quat(rot), 1, 2)
    return _positive_w(quat_from_rot(torch.bmm(inv, rot_from_quat(orient))))


def tf_from_pose(t: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """math.py:85-100."""
    m = torch.zeros(q.shape[0], 4, 4, dtype=torch.float32)
    m[:, :3, :3] = rot_from_quat(q)
    m[:, :3, 3] = t
    m[:, 3, 3] = 1
    return m


def tf_inverse(m: torch.Tensor) -> torch.Tensor:
    """math.py:152-173."""
    n = m.shape[0]
    rt = torch.transpose(m[:, 0:3, 0:3], 1, 2)
    p = torch.bmm(rt, (-m[:, 0:3, 3]).view(n, 3, 1))
    out = torch.zeros(n, 4, 4, dtype=torch.float32)
    out[:, 0:3, 0:3] = rt
    out[:, 0:3, 3] = p.view(n, 3)
    out[:, 3, 3] = 1.0
    return out


def tf_apply(m: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """math.py:129-150 — homogeneous [N,4,m] bmm."""
    vt = torch.transpose(v, 1, 2)
    h = torch.ones(vt.shape[0], 4, vt.shape[2], dtype=torch.float32)
    h[:, 0:3, :] = vt
    return torch.transpose(torch.bmm(m, h)[:, 0:3, :], 1, 2)


def transform_vectors(q, t, v):
    """math.py:102-113."""
    return tf_apply(tf_from_pose(t, q), v)


def inverse_transform_vectors(q, t, v):
    """math.py:115-127."""
    return tf_apply(tf_inverse(tf_from_pose(t, q)), v)


def rand_quaternions_from_uniforms(u: torch.Tensor, cfg: TaskConfig) -> torch.Tensor:
    """math.py:176-193 with the ``torch.rand((k,3))`` draw supplied by the caller."""
    roll = cfg.min_roll + (cfg.max_roll - cfg.min_roll) * u[:, 0]
    pitch = cfg.min_pitch + (cfg.max_pitch - cfg.min_pitch) * u[:, 1]
    yaw = cfg.min_yaw + (cfg.max_yaw - cfg.min_yaw) * u[:, 2]
    return it.quat_from_euler_xyz(roll, pitch, yaw)


# ----------------------------------------------------------------------------- the task
class TaskPort:
    """Restated task over a :class:`SyntheticScene` living on the CPU."""

    def __init__(self, cfg: TaskConfig, scene):
        self.cfg = cfg
        self.scene = scene
        n = self.num_envs = scene.num_envs
        self.half = scene.num_scenes
        self.has_loco = cfg.kind in (KIND_LOCO, KIND_COTRAIN)
        self.has_mani = cfg.kind in (KIND_MANI, KIND_COTRAIN)
        self.env_pos = scene.env_pos
        f32, i64 = torch.float32, torch.long
        # rl_task.py:104-113
        self.obs_buf = torch.zeros(n, cfg.num_observations, dtype=f32)
        self.states_buf = torch.zeros(n, cfg.num_states, dtype=f32)
        self.rew_buf = torch.zeros(n, dtype=f32)
        self.reset_buf = torch.ones(n, dtype=i64)
        self.progress_buf = torch.zeros(n, dtype=i64)
        self.extras: Dict[str, torch.Tensor] = {}
        # QPC:123-154
        self.last_actions = torch.zeros(n, 12, dtype=f32)
        self.current_actions = torch.zeros(n, 12, dtype=f32)
        self.last_base_tip_positions = torch.zeros(n, 4, 3, dtype=f32)
        self.default_base_tip_positions = torch.tensor(DEFAULT_BASE_TIP_POSITIONS, dtype=f32).repeat(n, 1, 1)
        self.goal_quaternions = torch.zeros(n, 4, dtype=f32)
        self.successes = torch.zeros(n, dtype=i64)
        self.consecutive_successes = torch.zeros(n, dtype=i64)
        self.goal_reset_buf = torch.zeros(n, dtype=i64)
        self.last_joint_velocities = {"loco": torch.zeros(self.half, 12, dtype=f32),
                                      "mani": torch.zeros(self.half, 12, dtype=f32)}   # robot.py:263
        self.targets = {"loco": torch.zeros(self.half, 12, dtype=f32), "mani": torch.zeros(self.half, 12, dtype=f32)}
        self.counters = {k: torch.tensor(0, dtype=i64) for k in
                         ("num_successes", "num_resets", "num_successes_loco", "num_resets_loco",
                          "num_successes_mani", "num_resets_mani")}
        self.rates = {k: torch.tensor(0.0) for k in ("success_rate", "success_rate_loco", "success_rate_mani")}
        h = self.half
        self.corners_robot = torch.tensor(cfg.corners, dtype=f32).repeat(h, 1, 1)           # QPC:132-137
        self.q0 = torch.tensor(cfg.mani_default_quaternion, dtype=f32).repeat(h, 1)          # QMP:149
        self.p0 = torch.tensor(cfg.mani_default_position, dtype=f32).repeat(h, 1)
        if self.has_mani:
            self.corners_world = transform_vectors(self.q0, self.p0, self.corners_robot)       # QMP:162-165
        self.time_interval = cfg.time_interval
        self.pd = cfg.control_mode == CONTROL_PD
        if self.pd:                                                                     # …position_control.py:179-191
            init = torch.zeros(n, 12, dtype=f32)
            li, mi = torch.tensor(cfg.loco_init_joint_pos[:12], dtype=f32), torch.tensor(cfg.mani_init_joint_pos[:12], dtype=f32)
            if cfg.kind == KIND_COTRAIN:
                init[:h], init[h:] = li, mi
            else:
                init[:] = li if cfg.kind == KIND_LOCO else mi
            self.default_joint_positions = init
            self.joint_positions = init.clone()
            self.joint_velocities = torch.zeros(n, 12, dtype=f32)
            self.current_joint_position_targets = init.clone()
            self.last_joint_position_targets = init.clone()
            self.current_joint_position_targets_se = torch.tensor(cfg.pd_init_se, dtype=f32).repeat(n, 1)
            self.se_upper = torch.tensor(cfg.pd_max_se, dtype=f32)
            self.se_lower = torch.tensor(cfg.pd_min_se, dtype=f32)
            self.torque = torch.zeros(n, 12, dtype=f32)
            self.applied_torque = torch.zeros(n, 12, dtype=f32)
        self.cc = bool(cfg.custom_controller)
        if self.cc and cfg.cc_rot_dist_term:
            self.last_rot_dist = torch.zeros(n, dtype=f32)                              # …custom_controller.py:188
        self.reset_env_ids = torch.zeros(0, dtype=i64)
        self.inter: Dict[str, torch.Tensor] = {}

    # ------------------------------------------------------------------ pre_physics_step  (QPC:200-228, JLV:248-270)
    def pre_physics_step(self, actions: torch.Tensor, uniforms: torch.Tensor, physics_substep=None) -> None:
        cfg = self.cfg
        self.current_actions = actions.clone()
        ids = self.reset_buf.clone().nonzero(as_tuple=False).squeeze(-1)
        self.reset_env_ids = ids
        if len(ids) > 0:
            self.reset_idx(ids, uniforms)
        act = torch.zeros_like(actions) if cfg.zero_actions else actions
        if self.pd:
            self._pd_loop(act, physics_substep)
            return
        h = self.half
        if self.has_loco:
            self.targets["loco"] = self._take_action(act[0:h])
        if self.has_mani:
            self.targets["mani"] = self._take_action(act[h:] if cfg.kind == KIND_COTRAIN else act)

    def _pd_state_from_scene(self):
        """update_joint_states of every robot (robot.py:283-291), concatenated locomotion-first."""
        jps, jvs = [], []
        for which in (("loco",) if self.has_loco else ()) + (("mani",) if self.has_mani else ()):
            jp, jv, _ = self._joint_state(which)
            jps.append(jp)
            jvs.append(jv)
        return torch.cat(jps, dim=0), torch.cat(jvs, dim=0)

    def _pd_torque(self) -> torch.Tensor:
        cfg = self.cfg
        torque = cfg.pd_kp * (self.current_joint_position_targets - self.joint_positions) - cfg.pd_kd * self.joint_velocities
        return torch.clamp(torque, min=-cfg.pd_max_effort, max=cfg.pd_max_effort)

    def _pd_loop(self, act: torch.Tensor, physics_substep) -> None:
        """quadruped_pose_control_position_control.py:275-313 / joint_locomanipulation_position_control.py:349-408."""
        cfg = self.cfg
        delta = act * cfg.pd_action_scale
        se = self.current_joint_position_targets_se + delta
        se = torch.clamp(se, min=self.se_lower, max=self.se_upper)
        self.current_joint_position_targets_se = se
        swing, ext = se[:, [4, 6, 8, 10]], se[:, [5, 7, 9, 11]]
        self.current_joint_position_targets[:, 0:4] = se[:, 0:4]
        self.current_joint_position_targets[:, [4, 6, 8, 10]] = swing + ext / 2.0
        self.current_joint_position_targets[:, [5, 7, 9, 11]] = swing - ext / 2.0
        torque_sum = 0.0
        for i in range(cfg.pd_control_decimal):
            torque = self._pd_torque()
            torque_sum += torque
            if physics_substep is not None:
                physics_substep(i)                                                       # SimulationContext.step
            self.joint_positions, self.joint_velocities = self._pd_state_from_scene()
        torque = self._pd_torque()
        if cfg.pd_sum_final_torque:                                                      # QPC_PC:311; absent in JL_PC:399-408
            torque_sum += torque
        self.applied_torque = torque
        self.torque = torque_sum / cfg.pd_torque_divisor

    def _take_action(self, a: torch.Tensor) -> torch.Tensor:
        """robot.py:444-461."""
        cfg = self.cfg
        lim = {CONTROL_VELOCITY: cfg.velocity_limit, CONTROL_POSITION: float(np.float32(np.pi)),
               CONTROL_EFFORT: cfg.torque_limit}[cfg.control_mode]
        hi = torch.full((12,), lim, dtype=torch.float32)
        return it.unscale_transform(a, -hi, hi)

    # ------------------------------------------------------------------ reset_idx  (QPC:230-299, QMP:238-309, JLV:272-377)
    def reset_idx(self, env_ids: torch.Tensor, uniforms: torch.Tensor) -> None:
        cfg, h = self.cfg, self.half
        k = len(env_ids)
        if cfg.kind == KIND_COTRAIN:
            ids_l = env_ids[env_ids < h]
            ids_m_off = env_ids[env_ids >= h]
            ids_m = ids_m_off - h
        elif cfg.kind == KIND_LOCO:
            ids_l, ids_m_off, ids_m = env_ids, env_ids[:0], env_ids[:0]
        else:
            ids_l, ids_m_off, ids_m = env_ids[:0], env_ids, env_ids
        kl, km = len(ids_l), len(ids_m)
        goal_l = rand_quaternions_from_uniforms(uniforms[0:kl], cfg)            # first draw (JLV:286-293)
        goal_m = rand_quaternions_from_uniforms(uniforms[kl:kl + km], cfg)      # second draw (JLV:294-301)
        f32 = torch.float32
        if kl > 0:
            s = self.scene.loco
            s["dof_pos"][ids_l] = torch.tensor(cfg.loco_init_joint_pos, dtype=f32).repeat(kl, 1)
            s["dof_vel"][ids_l] = torch.zeros(kl, 20, dtype=f32)
            s["root_pos"][ids_l] = torch.tensor(cfg.loco_default_position, dtype=f32).repeat(kl, 1) + self.env_pos[ids_l]
            s["root_quat"][ids_l] = torch.tensor(cfg.loco_default_quaternion, dtype=f32).repeat(kl, 1)
            s["root_lin_vel"][ids_l] = 0.0
            s["root_ang_vel"][ids_l] = 0.0
            self.last_joint_velocities["loco"][ids_l] = 0.0
            self.goal_quaternions[ids_l] = goal_l
            s["indicator_quat"][ids_l] = it.quat_conjugate(goal_l)
            s["indicator_pos"][ids_l] = self.env_pos[ids_l] + torch.tensor(cfg.indicator_loco_position, dtype=f32).repeat(kl, 1)
        if km > 0:
            s = self.scene.mani
            s["dof_pos"][ids_m] = torch.tensor(cfg.mani_init_joint_pos, dtype=f32).repeat(km, 1)
            s["dof_vel"][ids_m] = torch.zeros(km, 20, dtype=f32)
            s["plate_pos"][ids_m] = self.env_pos[ids_m] + torch.tensor(cfg.obj_default_position, dtype=f32).repeat(km, 1)
            s["plate_quat"][ids_m] = self.q0[ids_m]
            s["plate_lin_vel"][ids_m] = 0.0
            s["plate_ang_vel"][ids_m] = 0.0
            self.last_joint_velocities["mani"][ids_m] = 0.0
            self.goal_quaternions[ids_m_off] = goal_m
            s["indicator_quat"][ids_m] = rotate_orientations(self.q0[ids_m], goal_m)              # QMP:291 (SciPy)
            s["indicator_pos"][ids_m] = self.env_pos[ids_m] + torch.tensor(cfg.indicator_mani_position, dtype=f32).repeat(km, 1)
        if self.pd:                                                                     # …position_control.py:377-386
            self.joint_positions[env_ids] = self.default_joint_positions[env_ids]
            self.joint_velocities[env_ids] = 0.0
            self.current_joint_position_targets[env_ids] = self.default_joint_positions[env_ids]
            self.last_joint_position_targets[env_ids] = self.current_joint_position_targets[env_ids].clone()
            self.current_joint_position_targets_se[env_ids] = torch.tensor(cfg.pd_init_se, dtype=f32).repeat(k, 1)
        if self.cc and cfg.cc_rot_dist_term:                                            # …custom_controller.py:384-386
            q_def = torch.tensor(cfg.loco_default_quaternion, dtype=f32).repeat(k, 1)
            qd = it.quat_mul(it.quat_conjugate(q_def), it.quat_conjugate(self.goal_quaternions[env_ids]))
            self.last_rot_dist[env_ids] = 2.0 * torch.asin(torch.clamp(torch.norm(qd[:, 1:4], p=2, dim=-1), max=1.0))
        self.last_base_tip_positions[env_ids] = self.default_base_tip_positions[env_ids]
        self.last_actions[env_ids] = 0.0
        for buf in (self.reset_buf, self.progress_buf, self.successes, self.consecutive_successes, self.goal_reset_buf):
            buf[env_ids] = 0

    # ------------------------------------------------------------------ post_physics_step  (rl_task.py:240-260)
    def post_physics_step(self):
        self.progress_buf[:] += 1
        self.get_observations()
        self.calculate_metrics()
        self.is_done()
        return self.obs_buf, self.rew_buf, self.reset_buf, self.extras

    def _joint_state(self, which: str):
        """robot.py:283-291 (actuated DOFs are columns 0..11 of the 20-wide rows)."""
        s = getattr(self.scene, which)
        jp = s["dof_pos"][:, 0:12]
        jv = s["dof_vel"][:, 0:12]
        acc = (jv - self.last_joint_velocities[which]) / self.time_interval
        self.last_joint_velocities[which] = jv.clone()
        return jp, jv, acc

    def _rows(self, which: str):
        if self.cfg.kind != KIND_COTRAIN:
            return slice(0, self.num_envs)
        return slice(0, self.half) if which == "loco" else slice(self.half, self.num_envs)

    def _observe_loco(self) -> Dict[str, torch.Tensor]:
        cfg, s, h = self.cfg, self.scene.loco, self.half
        rows = self._rows("loco")
        jp, jv, acc = self._joint_state("loco")
        base_q = s["root_quat"]
        base_p = s["root_pos"] - self.env_pos                                       # robot.py:293-295
        tips = s["tip_pos"].view(h, 4, 3) - self.env_pos.repeat(1, 4).view(h, 4, 3)  # robot.py:305-310
        knees = s["knee_pos"].view(h, 8, 3) - self.env_pos.repeat(1, 8).view(h, 8, 3)
        pos = it.quat_rotate_inverse(base_q, -base_p)                               # QPC:319
        cur_q = it.quat_conjugate(base_q)                                           # QPC:322
        lin = it.quat_rotate_inverse(base_q, -s["root_lin_vel"])
        ang = it.quat_rotate_inverse(base_q, -s["root_ang_vel"])
        tip_base = inverse_transform_vectors(base_q, base_p, tips)                  # QPC:332-335
        corner_z = transform_vectors(base_q, base_p, self.corners_robot)[:, :, 2].view(h, 4)   # QPC:574-578
        return dict(rows=rows, jp=jp, jv=jv, acc=acc, pos=pos, cur_q=cur_q, lin=lin, ang=ang, tip_base=tip_base,
                    body_height=base_p[:, 2], corner_z=corner_z, knee_z=knees[:, :, 2])

    def _observe_mani(self) -> Dict[str, torch.Tensor]:
        cfg, s, h = self.cfg, self.scene.mani, self.half
        rows = self._rows("mani")
        jp, jv, acc = self._joint_state("mani")
        tips = s["tip_pos"].view(h, 4, 3) - self.env_pos.repeat(1, 4).view(h, 4, 3)
        knees = s["knee_pos"].view(h, 8, 3) - self.env_pos.repeat(1, 8).view(h, 8, 3)
        plate_p = s["plate_pos"] - self.env_pos                                      # rigid_object.py:52-55
        plate_q = s["plate_quat"]
        pos = inverse_transform_vectors(self.q0, self.p0, plate_p.view(h, 1, 3)).view(h, 3)     # QMP:326-329
        cur_q = inverse_rotate_orientations(self.q0, plate_q)                        # QMP:332
        lin = it.quat_rotate_inverse(self.q0, s["plate_lin_vel"])                    # QMP:338 (not negated)
        ang = it.quat_rotate_inverse(self.q0, s["plate_ang_vel"])
        tip_base = inverse_transform_vectors(self.q0, self.p0, tips)                 # QMP:343
        body = inverse_transform_vectors(plate_q, plate_p, self.p0.view(h, 1, 3)).view(h, 3)[:, 2]   # QMP:576-580
        corner_z = inverse_transform_vectors(plate_q, plate_p, self.corners_world)[:, :, 2].view(h, 4)
        knee_z = inverse_transform_vectors(plate_q, plate_p, knees)[:, :, 2].view(h, 8)
        return dict(rows=rows, jp=jp, jv=jv, acc=acc, pos=pos, cur_q=cur_q, lin=lin, ang=ang, tip_base=tip_base,
                    body_height=body, corner_z=corner_z, knee_z=knee_z)

    def get_observations(self) -> None:
        cfg = self.cfg
        parts = []
        if self.has_loco:
            parts.append(self._observe_loco())
        if self.has_mani:
            parts.append(self._observe_mani())
        obs_rows, state_rows = [], []
        for p in parts:
            rows = p["rows"]
            goal = self.goal_quaternions[rows]
            diff = it.quat_mul(p["cur_q"], it.quat_conjugate(goal))
            if not self.cc:                                                           # …custom_controller.py:430 keeps the raw product
                diff = _positive_w(diff)                                              # QPC:345-349
            up = it.quat_axis(p["cur_q"], 2)
            s_pos = cfg.position_scale * p["pos"]
            s_lin = cfg.linear_vel_scale * p["lin"]
            s_ang = cfg.angular_vel_scale * p["ang"]
            s_jp = cfg.joint_position_scale * p["jp"]
            s_jv = cfg.joint_velocity_scale * p["jv"]
            cur_a, last_a = self.current_actions[rows], self.last_actions[rows]
            if self.pd and not self.cc:                                               # …position_control.py:438-452
                cur_a = (self.current_joint_position_targets * cfg.pd_target_obs_scale)[rows]
                last_a = (self.last_joint_position_targets * cfg.pd_target_obs_scale)[rows]
            row = (s_pos, up, diff, s_lin, s_ang, s_jp, s_jv, cur_a, last_a)                                   # QPC:358-371
            if self.cc:                                                               # …custom_controller.py:432-455: actions AND target rows
                row = row + ((self.current_joint_position_targets * cfg.pd_target_obs_scale)[rows],
                             (self.last_joint_position_targets * cfg.pd_target_obs_scale)[rows])
            obs_rows.append(torch.cat(row, dim=-1))
            if cfg.num_states == 93:                                                  # QPC:406-423
                n = p["jp"].shape[0]
                tip_now = cfg.tip_position_scale * p["tip_base"].reshape(n, 12)
                tip_last = cfg.tip_position_scale * self.last_base_tip_positions[rows].clone().reshape(n, 12)
                state_rows.append(torch.cat((s_pos, s_lin, cfg.quaternion_obs_scale * p["cur_q"], s_ang, s_jp, s_jv,
                                             cfg.quaternion_obs_scale * goal, diff, tip_now, tip_last, cur_a, last_a), dim=-1))
        self.obs_buf[:] = torch.cat(obs_rows, dim=0)
        if cfg.num_states == 93:
            self.states_buf[:] = torch.cat(state_rows, dim=0)
        else:
            self.states_buf = self.obs_buf                                            # JLV:515
        if self.pd and cfg.pd_carry_post_state:                                       # QPC_PC:392-393
            self.joint_positions = torch.cat([p["jp"] for p in parts], dim=0)
            self.joint_velocities = torch.cat([p["jv"] for p in parts], dim=0)
        self.last_base_tip_positions = torch.cat([p["tip_base"] for p in parts], dim=0)   # QPC:426 / JLV:518-521
        cat = lambda k: torch.cat([p[k] for p in parts], dim=0)
        self.inter = {k: cat(k) for k in ("jp", "acc", "pos", "cur_q", "body_height", "corner_z", "knee_z")}

    # ------------------------------------------------------------------ calculate_metrics  (QPC:428-560, JLV:523-653)
    def calculate_metrics(self) -> None:
        cfg, x = self.cfg, self.inter
        quat_diff = it.quat_mul(x["cur_q"], it.quat_conjugate(self.goal_quaternions))
        rot_dist = 2.0 * torch.asin(torch.clamp(torch.norm(quat_diff[:, 1:4], p=2, dim=-1), max=1.0))
        rot_rew = 1.0 / (torch.abs(rot_dist) + cfg.rot_eps) * cfg.quaternion_scale
        translation_penalty = torch.norm(x["pos"][:, 0:2], dim=-1) * cfg.translation_scale
        joint_acc_penalty = torch.sum(torch.abs(x["acc"]) * cfg.joint_acc_scale, dim=1)
        if self.cc:                                                                   # …custom_controller.py:531
            action_rate_penalty = torch.sum(torch.abs(self.current_actions), dim=1) * cfg.action_rate_scale
        else:
            action_rate_penalty = torch.sum(torch.abs(self.last_actions - self.current_actions), dim=1) * cfg.action_rate_scale
        goal_reset = torch.where(self.consecutive_successes > cfg.max_consecutive_successes,
                                 torch.ones_like(self.consecutive_successes), torch.zeros_like(self.consecutive_successes))
        bonus = cfg.consecutive_success_reward * goal_reset.to(torch.float32)
        successes = torch.where(torch.abs(rot_dist) <= cfg.success_thresh, torch.ones_like(goal_reset), torch.zeros_like(goal_reset))
        jp = x["jp"]
        d23 = torch.abs(jp[:, [5, 7, 9, 11]] - jp[:, [4, 6, 8, 10]])
        brk23 = torch.sum((d23 < cfg.min_joint_23_diff) + (d23 > cfg.max_joint_23_diff), dim=-1)
        self.joint23_pos_reset = torch.sum((d23 < cfg.reset_min_joint_23_diff) + (d23 > cfg.reset_max_joint_23_diff), dim=-1)
        if cfg.joint1_mirrored:                                                       # QPC:466-504
            a14, a23 = jp[:, [0, 3]], jp[:, [1, 2]]
            brk1 = torch.sum((a14 < cfg.min_joint_1_pos) + (a14 > cfg.max_joint_1_pos)
                             + (a23 < -cfg.max_joint_1_pos) + (a23 > -cfg.min_joint_1_pos), dim=-1)
            self.joint1_pos_reset = torch.sum((a14 < cfg.reset_min_joint_1_pos) + (a14 > cfg.reset_max_joint_1_pos)
                                              + (a23 < -cfg.reset_max_joint_1_pos) + (a23 > -cfg.reset_min_joint_1_pos), dim=-1)
        else:                                                                         # JLV:574-597
            d1 = jp[:, 0:4]
            brk1 = torch.sum((d1 < cfg.min_joint_1_pos) + (d1 > cfg.max_joint_1_pos), dim=-1)
            self.joint1_pos_reset = torch.sum((d1 < cfg.reset_min_joint_1_pos) + (d1 > cfg.reset_max_joint_1_pos), dim=-1)
        brk = (brk1 + brk23).flatten()
        joint_limit_penalty = (cfg.joint_limit_penalty * torch.where(brk > 0, torch.ones_like(brk), torch.zeros_like(brk))).to(torch.float32)
        total = rot_rew + translation_penalty + joint_acc_penalty + action_rate_penalty + bonus + joint_limit_penalty
        extra_terms = {}
        if self.cc:                                                                   # …custom_controller.py:533-546,608-620
            mechanical_power = torch.sum(torch.abs(self.torque * self.joint_velocities), dim=-1)
            mechanical_power_penalty = cfg.mechanical_power_penalty_scale * mechanical_power
            position_target_error = torch.sum(torch.abs(self.last_joint_position_targets - self.joint_positions), dim=-1)
            position_target_error_penalty = cfg.position_target_error_penalty_scale * position_target_error
            total = total + mechanical_power_penalty + position_target_error_penalty
            extra_terms = {"mechanical_power_penalty": mechanical_power_penalty, "position_target_error_penalty": position_target_error_penalty}
            if cfg.cc_rot_dist_term:                                                  # locomotion file only (:541-546)
                larger = rot_dist > cfg.no_rot_dist_decreasing_reward_thresh
                rot_dist_decreasing_reward = larger.to(torch.float32) * (self.last_rot_dist - rot_dist) * cfg.rot_dist_decreasing_reward_scale
                self.last_rot_dist = rot_dist
                total = total + rot_dist_decreasing_reward
                extra_terms["rot_dist_decreasing_reward"] = rot_dist_decreasing_reward
            if cfg.clip_reward:
                total = torch.clip(total, 0.0, None)
        self.rew_buf[:] = total
        self.goal_reset_buf = goal_reset
        both = torch.logical_and(successes, self.successes)                           # QPC:534-545
        cs = self.consecutive_successes
        cs[:] = torch.where(both == 0, torch.zeros_like(cs), cs)
        cs[:] = torch.where(both == 1, cs + 1, cs)
        first = torch.logical_and(self.successes == 0, successes == 1)
        cs[:] = torch.where(first == 1, torch.ones_like(cs), cs)
        self.successes[:] = successes
        self.last_actions = self.current_actions.clone()
        terms = {"orientation_rew": rot_rew, "translation_penalty": translation_penalty, "joint_acc_penalty": joint_acc_penalty,
                 "action_rate_penalty": action_rate_penalty, "consecutive_successes_rew": bonus,
                 "joint_limit_panelty": joint_limit_penalty}
        terms.update(extra_terms)
        self.extras.update({"env/rewards/" + k: v.mean() for k, v in terms.items()})
        self.inter["rot_dist"] = rot_dist
        self.inter["quat_diff_norm"] = torch.norm(quat_diff[:, 1:4], p=2, dim=-1)
        self.inter["d23"] = d23
        # magnitude of the summed terms: the natural scale for a relative tolerance on a reward that cancels to ~0
        self.inter["rew_scale"] = sum(v.abs() for v in terms.values())

    # ------------------------------------------------------------------ is_done  (QPC:562-633, JLV:655-802)
    def is_done(self) -> None:
        cfg, x, rb = self.cfg, self.inter, self.reset_buf
        one = torch.ones_like(rb)
        rb[:] = torch.where(x["pos"][:, 2] > 0.0, one, rb)
        rb[:] = torch.where(x["body_height"] <= cfg.baseline_height, one, rb)
        rb[:] = torch.where(torch.sum(x["corner_z"] < cfg.baseline_corner_height, dim=-1) > 0, one, rb)
        rb[:] = torch.where(torch.sum(x["knee_z"] - cfg.baseline_knee_height <= 0, dim=1) > 0, one, rb)
        rb[:] = torch.where(self.joint1_pos_reset + self.joint23_pos_reset > 0, one, rb)
        fall = (cfg.fall_penalty * rb).to(torch.float32)
        self.rew_buf += fall
        self.extras["env/rewards/fall_penalty"] = fall.mean()
        rb[:] = torch.where(self.goal_reset_buf == 1, one, rb)
        rb[:] = torch.where(self.progress_buf >= cfg.max_episode_length - 1, one, rb)
        if self.cc and cfg.cc_update_last_targets:                                    # …custom_controller.py:724
            self.last_joint_position_targets = self.current_joint_position_targets.clone()
        groups = [("", slice(0, self.num_envs))]
        if cfg.kind == KIND_COTRAIN:                                                   # JLV:769-798
            groups += [("_loco", slice(0, self.half)), ("_mani", slice(self.half, self.num_envs))]
        c, r = self.counters, self.rates
        for sfx, _ in groups:                                                          # QPC:619-628
            if c["num_resets" + sfx] > cfg.max_reset_counts:
                r["success_rate" + sfx] = c["num_successes" + sfx] / c["num_resets" + sfx]
                c["num_resets" + sfx] = torch.tensor(0, dtype=torch.long)
                c["num_successes" + sfx] = torch.tensor(0, dtype=torch.long)
        for sfx, rows in groups:
            c["num_successes" + sfx] = c["num_successes" + sfx] + torch.sum(self.goal_reset_buf[rows])
            c["num_resets" + sfx] = c["num_resets" + sfx] + torch.sum(rb[rows])
        for sfx, _ in groups:
            self.extras["env/success_rate" + sfx] = r["success_rate" + sfx]

    # ------------------------------------------------------------------ helpers for the parity tests
    def threshold_margins(self) -> torch.Tensor:
        """Per-env min |value - threshold| over every float compare that feeds a mask (last step).
        Tests exempt an env from the bit-exact mask check only if this margin is below a few ulp."""
        cfg, x = self.cfg, self.inter
        m = torch.abs(torch.abs(x["rot_dist"]) - cfg.success_thresh)
        m = torch.minimum(m, torch.abs(x["pos"][:, 2]))
        m = torch.minimum(m, torch.abs(x["body_height"] - cfg.baseline_height))
        m = torch.minimum(m, torch.abs(x["corner_z"] - cfg.baseline_corner_height).min(dim=-1).values)
        m = torch.minimum(m, torch.abs(x["knee_z"] - cfg.baseline_knee_height).min(dim=-1).values)
        return m
