"""Synthetic PhysX stand-in: articulation / rigid-body state tensors at the tasks' real shapes.

Rigid-body stepping is out of scope (BASELINE.json north_star); this module plays the role of the
``ArticulationView`` / ``RigidPrimView`` tensor API the reference reads through
(robot/base/robot.py:288-321, objects/base/rigid_object.py:52-58) and writes through on reset
(robot.py:357-425, rigid_object.py:36-50).  State is seeded per step (``seed + step``) following
SURVEY.md §8(d): joint pose = default + U(-a, a), small base wobble, tips near the default tip
positions, a few % of envs tripping each termination test, and a fixed subset of envs "tracking"
their goal so the consecutive-success / goal-reset state machine is exercised.

All generation happens with a CPU ``torch.Generator`` (identical streams on every box) and is then
moved to the target device; nothing here is on the timed path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from .configs import (DEFAULT_BASE_TIP_POSITIONS, KIND_COTRAIN, KIND_LOCO, KIND_MANI, TaskConfig)

NUM_DOF = 20          # 12 actuated + 8 passive (robot.py:437, quadruped_robot.py:45-52)
NUM_ACTUATED = 12
NUM_MODULES = 4


# ----------------------------------------------------------------------------- small quaternion helpers (wxyz)
def _quat_from_euler(roll, pitch, yaw):
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    return torch.stack([cy * cr * cp + sy * sr * sp, cy * sr * cp - sy * cr * sp,
                        cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp], dim=-1)


def _quat_mul(a, b):
    w1, x1, y1, z1 = a.unbind(-1)
    w2, x2, y2, z2 = b.unbind(-1)
    return torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
                        w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                        w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                        w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], dim=-1)


def _quat_conj(q):
    return torch.cat([q[..., :1], -q[..., 1:]], dim=-1)


def _quat_rotate(q, v):
    """Rotate v[..., 3] by unit quaternion q[..., 4] (wxyz)."""
    w = q[..., :1]
    u = q[..., 1:]
    t = 2.0 * torch.cross(u, v, dim=-1)
    return v + w * t + torch.cross(u, t, dim=-1)


def _normalize(q):
    return q / q.norm(dim=-1, keepdim=True)


def grid_env_origins(num_scenes: int, spacing: float) -> torch.Tensor:
    """Env origins on a square grid (the role of GridCloner.clone, rl_task.py:129-131)."""
    per_row = max(1, int(math.ceil(math.sqrt(num_scenes))))
    idx = torch.arange(num_scenes)
    xy = torch.stack([(idx % per_row).float(), (idx // per_row).float()], dim=-1) * spacing
    xy = xy - (per_row - 1) * spacing * 0.5
    return torch.cat([xy, torch.zeros(num_scenes, 1)], dim=-1).contiguous()


@dataclass
class SynthKnobs:
    joint_spread: float = 0.34        # U(-a, a) on the 12 actuated DOFs (capped so |dof3-dof2| rarely leaves its band)
    joint_vel_std: float = 0.5
    base_xy_std: float = 0.03
    base_z_std: float = 0.02
    base_rp_range: float = 0.385      # roll / pitch of the base, U(±)
    base_yaw_range: float = 1.6
    lin_vel_std: float = 0.2
    ang_vel_std: float = 0.5
    tip_noise_std: float = 0.01
    knee_low: float = 0.05
    knee_high: float = 0.12
    p_low_knee: float = 0.006         # per env: one knee dropped below the 0.04 threshold
    p_fallen: float = 0.004           # per env: body height dropped below 0.05
    plate_xy_std: float = 0.02
    plate_z_std: float = 0.01
    plate_tilt_range: float = 0.35
    track_every: int = 8              # env e tracks its goal iff e % track_every == 0 (0 disables)
    track_noise: float = 0.03         # small-angle noise (rad) of tracking envs around the goal


def _articulation_joints(gen, n, init_joint_pos, knobs):
    default = torch.tensor(init_joint_pos, dtype=torch.float32).repeat(n, 1)
    jp = default.clone()
    # keep |dof3 - dof2| mostly inside [0.43, 2.53] (QPC:80-81) so the joint-limit tests fire for ~1 % of envs
    diff0 = abs(init_joint_pos[5] - init_joint_pos[4])
    spread = min(knobs.joint_spread, max(0.02, (diff0 - 0.41) * 0.5))
    jp[:, :NUM_ACTUATED] += (torch.rand(n, NUM_ACTUATED, generator=gen) * 2.0 - 1.0) * spread
    jp[:, NUM_ACTUATED:] += torch.randn(n, NUM_DOF - NUM_ACTUATED, generator=gen) * 0.05
    jv = torch.randn(n, NUM_DOF, generator=gen) * knobs.joint_vel_std
    return jp.contiguous(), jv.contiguous()


def _small_rotation(gen, n, rp, yaw):
    r = (torch.rand(n, generator=gen) * 2 - 1) * rp
    p = (torch.rand(n, generator=gen) * 2 - 1) * rp
    y = (torch.rand(n, generator=gen) * 2 - 1) * yaw
    return _normalize(_quat_from_euler(r, p, y))


def synth_loco_state(cfg: TaskConfig, n: int, env_pos: torch.Tensor, gen: torch.Generator,
                     knobs: SynthKnobs, goal_quat: Optional[torch.Tensor] = None,
                     env_id_offset: int = 0) -> Dict[str, torch.Tensor]:
    """State of the free-base locomotion robot: what robot.py:276-321 reads each step."""
    jp, jv = _articulation_joints(gen, n, cfg.loco_init_joint_pos, knobs)
    z0 = cfg.loco_default_position[2]
    local = torch.stack([torch.randn(n, generator=gen) * knobs.base_xy_std,
                         torch.randn(n, generator=gen) * knobs.base_xy_std,
                         z0 + torch.randn(n, generator=gen) * knobs.base_z_std], dim=-1)
    fallen = torch.rand(n, generator=gen) < knobs.p_fallen
    fall_z = 0.02 + 0.03 * torch.rand(n, generator=gen)
    local[:, 2] = torch.where(fallen, fall_z, local[:, 2])
    quat = _small_rotation(gen, n, knobs.base_rp_range, knobs.base_yaw_range)
    noise_q = _small_rotation(gen, n, knobs.track_noise, knobs.track_noise)
    if goal_quat is not None and knobs.track_every > 0:
        # success needs ground_quat = conj(base_quat) ≈ goal (QPC:322,345)  →  base_quat ≈ conj(goal)
        ids = torch.arange(n) + env_id_offset
        track = (ids % knobs.track_every == 0) & (goal_quat.abs().sum(-1) > 0)
        tracked = _normalize(_quat_mul(_quat_conj(goal_quat.cpu().float()), noise_q))
        quat = torch.where(track[:, None], tracked, quat)
    lin = torch.randn(n, 3, generator=gen) * knobs.lin_vel_std
    ang = torch.randn(n, 3, generator=gen) * knobs.ang_vel_std
    tips_base = torch.tensor(DEFAULT_BASE_TIP_POSITIONS, dtype=torch.float32).repeat(n, 1, 1)
    tips_local = _quat_rotate(quat[:, None, :].expand(n, NUM_MODULES, 4), tips_base) + local[:, None, :]
    tips_local = tips_local + torch.randn(n, NUM_MODULES, 3, generator=gen) * knobs.tip_noise_std
    knees_local = torch.empty(n, 2 * NUM_MODULES, 3)
    knees_local[..., :2] = local[:, None, :2] + torch.randn(n, 2 * NUM_MODULES, 2, generator=gen) * 0.1
    knees_local[..., 2] = knobs.knee_low + (knobs.knee_high - knobs.knee_low) * torch.rand(n, 2 * NUM_MODULES, generator=gen)
    low = torch.rand(n, generator=gen) < knobs.p_low_knee
    which = torch.randint(0, 2 * NUM_MODULES, (n,), generator=gen)
    low_z = 0.02 + 0.02 * torch.rand(n, generator=gen)
    rows = torch.nonzero(low).squeeze(-1)
    knees_local[rows, which[rows], 2] = low_z[rows]
    origin = env_pos.cpu().float()
    return {
        "dof_pos": jp, "dof_vel": jv,
        "root_pos": (local + origin).contiguous(), "root_quat": quat.contiguous(),
        "root_lin_vel": lin.contiguous(), "root_ang_vel": ang.contiguous(),
        "tip_pos": (tips_local + origin[:, None, :]).reshape(n * NUM_MODULES, 3).contiguous(),
        "knee_pos": (knees_local + origin[:, None, :]).reshape(n * 2 * NUM_MODULES, 3).contiguous(),
    }


def synth_mani_state(cfg: TaskConfig, n: int, env_pos: torch.Tensor, gen: torch.Generator,
                     knobs: SynthKnobs, goal_quat: Optional[torch.Tensor] = None,
                     env_id_offset: int = 0) -> Dict[str, torch.Tensor]:
    """State of the inverted fixed-base robot + plate: what QMP:311-343 reads each step."""
    jp, jv = _articulation_joints(gen, n, cfg.mani_init_joint_pos, knobs)
    q0 = torch.tensor(cfg.mani_default_quaternion, dtype=torch.float32).repeat(n, 1)
    p0 = torch.tensor(cfg.mani_default_position, dtype=torch.float32).repeat(n, 1)
    pz = cfg.obj_default_position[2]
    plate_local = torch.stack([torch.randn(n, generator=gen) * knobs.plate_xy_std,
                               torch.randn(n, generator=gen) * knobs.plate_xy_std,
                               pz + torch.randn(n, generator=gen) * knobs.plate_z_std], dim=-1)
    rel = _small_rotation(gen, n, knobs.plate_tilt_range, knobs.base_yaw_range)
    noise_q = _small_rotation(gen, n, knobs.track_noise, knobs.track_noise)
    if goal_quat is not None and knobs.track_every > 0:
        # success needs plate_quat_robot = q0^-1 ⊗ plate_quat ≈ goal (QMP:332,354)
        ids = torch.arange(n) + env_id_offset
        track = (ids % knobs.track_every == 0) & (goal_quat.abs().sum(-1) > 0)
        rel = torch.where(track[:, None], _normalize(_quat_mul(goal_quat.cpu().float(), noise_q)), rel)
    plate_quat = _normalize(_quat_mul(q0, rel))
    lin = torch.randn(n, 3, generator=gen) * knobs.lin_vel_std
    ang = torch.randn(n, 3, generator=gen) * knobs.ang_vel_std
    tips_base = torch.tensor(DEFAULT_BASE_TIP_POSITIONS, dtype=torch.float32).repeat(n, 1, 1)
    tips_local = _quat_rotate(q0[:, None, :].expand(n, NUM_MODULES, 4), tips_base) + p0[:, None, :]
    tips_local = tips_local + torch.randn(n, NUM_MODULES, 3, generator=gen) * knobs.tip_noise_std
    # knees are tested in the plate frame (QMP:596-603): place them a plate-frame height in [low, high] under the plate
    h = knobs.knee_low + (knobs.knee_high - knobs.knee_low) * torch.rand(n, 2 * NUM_MODULES, generator=gen)
    low = torch.rand(n, generator=gen) < knobs.p_low_knee
    which = torch.randint(0, 2 * NUM_MODULES, (n,), generator=gen)
    low_h = 0.02 + 0.02 * torch.rand(n, generator=gen)
    rows = torch.nonzero(low).squeeze(-1)
    h[rows, which[rows]] = low_h[rows]
    knee_plate = torch.cat([torch.randn(n, 2 * NUM_MODULES, 2, generator=gen) * 0.1, h[..., None]], dim=-1)
    knees_local = _quat_rotate(plate_quat[:, None, :].expand(n, 2 * NUM_MODULES, 4), knee_plate) + plate_local[:, None, :]
    # "fallen": lift the plate frame so the robot origin sits < 0.05 above it (QMP:576-581)
    fallen = torch.rand(n, generator=gen) < knobs.p_fallen
    origin = env_pos.cpu().float()
    root_pos = p0 + origin
    state = {
        "dof_pos": jp, "dof_vel": jv,
        "root_pos": root_pos.contiguous(), "root_quat": q0.contiguous(),
        "root_lin_vel": torch.zeros(n, 3), "root_ang_vel": torch.zeros(n, 3),
        "tip_pos": (tips_local + origin[:, None, :]).reshape(n * NUM_MODULES, 3).contiguous(),
        "knee_pos": (knees_local + origin[:, None, :]).reshape(n * 2 * NUM_MODULES, 3).contiguous(),
        "plate_pos": None, "plate_quat": plate_quat.contiguous(),
        "plate_lin_vel": lin.contiguous(), "plate_ang_vel": ang.contiguous(),
    }
    # robot origin in plate frame has height (R_plate^T (p0 - plate))_z; move the plate along its own z axis
    zaxis = _quat_rotate(plate_quat, torch.tensor([0.0, 0.0, 1.0]).repeat(n, 1))
    cur_h = ((p0 - plate_local) * zaxis).sum(-1)
    want_h = 0.02 + 0.03 * torch.rand(n, generator=gen)
    shift = (cur_h - want_h)[:, None] * zaxis
    plate_local = torch.where(fallen[:, None], plate_local + shift, plate_local)
    state["plate_pos"] = (plate_local + origin).contiguous()
    return state


class SyntheticScene:
    """Holds the per-step synthetic state for one task instance, on ``device``.

    Tensors are allocated once and refilled in place by :meth:`step`, so pointers handed to the
    CUDA kernels stay valid (the reference reads PhysX buffers the same way: ``clone=False`` views,
    robot.py:288-303).
    """

    def __init__(self, cfg: TaskConfig, num_envs: int, device="cpu", seed: int = 42,
                 knobs: Optional[SynthKnobs] = None, env_id_offset: int = 0):
        self.cfg = cfg
        self.num_envs = num_envs
        self.device = torch.device(device)
        self.seed = seed
        self.knobs = knobs or SynthKnobs()
        self.env_id_offset = env_id_offset
        if cfg.kind == KIND_COTRAIN:
            assert num_envs % 2 == 0, "Number of envs must be a multiplier of 2."  # JLV:135-136
            self.num_scenes = num_envs // 2          # joint_rl_task.py:133
        else:
            self.num_scenes = num_envs
        self.env_pos = grid_env_origins(self.num_scenes, cfg.env_spacing).to(self.device)
        self.loco: Dict[str, torch.Tensor] = {}
        self.mani: Dict[str, torch.Tensor] = {}
        self.arena: Dict[str, torch.Tensor] = {}        # one flat fp32 buffer per robot: all view tensors are slices of it
        self.physics_floats: Dict[str, int] = {}        # leading floats of the arena that "physics" rewrites every step
        n = self.num_scenes
        if cfg.kind in (KIND_LOCO, KIND_COTRAIN):
            self.loco = self._carve("loco", self._art_shapes(n), {"indicator_pos": (n, 3), "indicator_quat": (n, 4)})
        if cfg.kind in (KIND_MANI, KIND_COTRAIN):
            shapes = dict(self._art_shapes(n))
            shapes.update({"plate_pos": (n, 3), "plate_quat": (n, 4), "plate_lin_vel": (n, 3), "plate_ang_vel": (n, 3)})
            self.mani = self._carve("mani", shapes, {"indicator_pos": (n, 3), "indicator_quat": (n, 4)})

    def _carve(self, group, physics_shapes, extra_shapes):
        """Allocate one flat buffer and slice 16-byte-aligned tensors out of it (so a step's state is ONE host->device copy)."""
        def pad4(x):
            return (x + 3) // 4 * 4
        offs, total = {}, 0
        for k, shp in list(physics_shapes.items()) + list(extra_shapes.items()):
            if k == next(iter(extra_shapes)):
                self.physics_floats[group] = total
            offs[k] = (total, shp)
            total += pad4(shp[0] * shp[1])
        flat = torch.zeros(total, device=self.device, dtype=torch.float32)
        self.arena[group] = flat
        return {k: flat[o:o + shp[0] * shp[1]].view(shp) for k, (o, shp) in offs.items()}

    @staticmethod
    def _art_shapes(n):
        return {"dof_pos": (n, NUM_DOF), "dof_vel": (n, NUM_DOF), "root_pos": (n, 3), "root_quat": (n, 4),
                "root_lin_vel": (n, 3), "root_ang_vel": (n, 3), "tip_pos": (n * NUM_MODULES, 3),
                "knee_pos": (n * 2 * NUM_MODULES, 3)}

    def generate(self, step: int, goal_quaternions: Optional[torch.Tensor] = None) -> Dict[str, Dict[str, torch.Tensor]]:
        """CPU tensors for ``step`` (no device copy)."""
        gen = torch.Generator(device="cpu")
        gen.manual_seed(self.seed + step)
        n = self.num_scenes
        out = {}
        env_pos = self.env_pos.cpu()
        gq = goal_quaternions.detach().cpu() if goal_quaternions is not None else None
        if self.loco:
            g = gq[:n] if gq is not None else None
            out["loco"] = synth_loco_state(self.cfg, n, env_pos, gen, self.knobs, g, self.env_id_offset)
        if self.mani:
            g = None
            if gq is not None:
                g = gq[n:] if self.cfg.kind == KIND_COTRAIN else gq[:n]
            out["mani"] = synth_mani_state(self.cfg, n, env_pos, gen, self.knobs, g, self.env_id_offset)
        return out

    def load(self, states: Dict[str, Dict[str, torch.Tensor]], non_blocking: bool = False) -> None:
        """Copy a generated (possibly pinned-host) state into the persistent device tensors."""
        for group, dst in (("loco", self.loco), ("mani", self.mani)):
            for k, v in states.get(group, {}).items():
                dst[k].copy_(v, non_blocking=non_blocking)

    def pack_host(self, states: Dict[str, Dict[str, torch.Tensor]], pin: bool = True) -> Dict[str, torch.Tensor]:
        """Lay a generated state out exactly like the device arenas (physics part only) in pinned host memory."""
        out = {}
        for group, views in (("loco", self.loco), ("mani", self.mani)):
            if not views:
                continue
            nfl = self.physics_floats[group]
            host = torch.zeros(nfl, dtype=torch.float32)
            base = self.arena[group].data_ptr()
            for k, v in states[group].items():
                off = (views[k].data_ptr() - base) // 4
                host[off:off + v.numel()] = v.reshape(-1)
            out[group] = host.pin_memory() if pin and torch.cuda.is_available() else host
        return out

    def substep_joint_state(self, step: int, i: int) -> Dict[str, Dict[str, torch.Tensor]]:
        """Joint rows after physics sub-step ``i`` of env step ``step`` (the explicit-PD variants re-read the joint state
        after every ``SimulationContext.step``, …position_control.py:300-305).  CPU tensors; only dof_pos / dof_vel."""
        gen = torch.Generator(device="cpu")
        gen.manual_seed(self.seed + 7919 * (step + 1) + 31 * (i + 1))
        out = {}
        if self.loco:
            jp, jv = _articulation_joints(gen, self.num_scenes, self.cfg.loco_init_joint_pos, self.knobs)
            out["loco"] = {"dof_pos": jp, "dof_vel": jv}
        if self.mani:
            jp, jv = _articulation_joints(gen, self.num_scenes, self.cfg.mani_init_joint_pos, self.knobs)
            out["mani"] = {"dof_pos": jp, "dof_vel": jv}
        return out

    def substep(self, step: int, i: int) -> None:
        self.load(self.substep_joint_state(step, i))

    def step(self, step: int, goal_quaternions: Optional[torch.Tensor] = None) -> None:
        self.load(self.generate(step, goal_quaternions))
