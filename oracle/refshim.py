"""TEST INFRASTRUCTURE — runs the reference's own task classes, unmodified, from /root/reference.

The reference is pure Python on top of Isaac Sim (absent here).  This module installs ~100 lines of
fake ``omni`` / ``pxr`` / ``carb`` / ``gym`` / ``torch_scatter`` modules (recipe: SURVEY.md Appendix A),
backs the PhysX tensor views with the product's :class:`SyntheticScene` tensors, and exposes
:func:`make_reference_task`.  It only works where /root/reference exists (the authoring container):
it is used to (1) validate the torch-CPU port in ``oracle/port.py`` and (2) generate the golden vectors
committed under tests/golden/ (script: oracle/make_golden.py).  Nothing in ``-m gpu`` tests, smoke()
or bench.py touches it.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from typing import Dict, List, Optional

import numpy as np
import torch

from . import isaac_torch

REFERENCE_ROOT = os.environ.get("LM_REFERENCE_ROOT", "/root/reference")
_RL_DIR = os.path.join(REFERENCE_ROOT, "RobotLearning")
_ENV_DIR = os.path.join(_RL_DIR, "omniisaacgymenvs")


def reference_available() -> bool:
    return os.path.isdir(_ENV_DIR)


# ----------------------------------------------------------------------------- fake modules
class _Anything:
    """Callable no-op object standing in for any simulator-side class / function."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _FakeModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


class _BaseTask:
    def __init__(self, name, offset=None):
        self._name = name
        self._offset = offset

    def set_up_scene(self, scene):
        pass


class _Robot:
    def __init__(self, *a, **k):
        pass


class _Box:
    def __init__(self, low, high, *a, **k):
        self.low, self.high = np.asarray(low), np.asarray(high)
        self.shape = self.low.shape


class _SimulationContext:
    """``SimulationContext.step(world, render=False)`` is called inside the *_PC PD loops
    (quadruped_pose_control_position_control.py:300-301); the harness hooks it to swap joint state."""
    hook = None

    @staticmethod
    def step(world=None, render=False):
        if _SimulationContext.hook is not None:
            _SimulationContext.hook()


def _scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    """torch_scatter.scatter(reduce='max') stand-in (graph_model_orebot_ov.py:4,46): zero-initialised
    output, segments reduced with amax (every node of the limb graph has >= 1 in-edge)."""
    assert reduce == "max"
    dim = dim if dim >= 0 else src.dim() + dim
    shape = list(src.shape)
    shape[dim] = int(dim_size)
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    view = [1] * src.dim()
    view[dim] = -1
    idx = index.view(view).expand_as(src)
    return res.scatter_reduce(dim, idx, src, reduce="amax", include_self=False)


_INSTALLED = False


def install() -> None:
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if not hasattr(np, "Inf"):
        np.Inf = np.inf  # numpy >= 2 removed it; used at tasks/base/rl_task.py:75-76,94,96
    names = [
        "omni", "omni.isaac", "omni.isaac.core", "omni.isaac.core.tasks", "omni.isaac.core.utils",
        "omni.isaac.core.utils.types", "omni.isaac.core.utils.prims", "omni.isaac.core.utils.stage",
        "omni.isaac.core.utils.torch", "omni.isaac.core.robots", "omni.isaac.core.robots.robot",
        "omni.isaac.core.articulations", "omni.isaac.core.prims", "omni.isaac.core.materials",
        "omni.isaac.core.simulation_context", "omni.isaac.cloner", "omni.isaac.gym", "omni.isaac.gym.vec_env",
        "omni.kit", "omni.kit.viewport", "omni.kit.viewport.utility", "omni.kit.viewport.utility.camera_state",
        "omni.replicator", "omni.replicator.core", "omni.replicator.isaac", "omni.usd", "pxr", "carb", "gym",
        "gym.spaces",
    ]
    for n in names:
        if n not in sys.modules:
            m = _FakeModule(n)
            m.__path__ = []  # behave as a package
            sys.modules[n] = m
    for n in names:  # wire children as attributes so ``import omni.kit`` then ``omni.kit.x`` works
        if "." in n:
            parent, child = n.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[n])
    sys.modules["omni.isaac.core.tasks"].BaseTask = _BaseTask
    sys.modules["omni.isaac.core.robots.robot"].Robot = _Robot
    sys.modules["omni.isaac.core.simulation_context"].SimulationContext = _SimulationContext
    sys.modules["gym.spaces"].Box = _Box
    sys.modules["gym"].spaces = sys.modules["gym.spaces"]
    rot = types.ModuleType("omni.isaac.core.utils.torch.rotations")
    for f in ("quat_mul", "quat_conjugate", "quat_rotate", "quat_rotate_inverse", "quat_axis", "quat_apply",
              "quat_from_euler_xyz"):
        setattr(rot, f, getattr(isaac_torch, f))
    maths = types.ModuleType("omni.isaac.core.utils.torch.maths")
    maths.unscale_transform = isaac_torch.unscale_transform
    maths.set_seed = isaac_torch.set_seed
    sys.modules["omni.isaac.core.utils.torch.rotations"] = rot
    sys.modules["omni.isaac.core.utils.torch.maths"] = maths
    sys.modules["omni.isaac.core.utils.torch"].rotations = rot
    sys.modules["omni.isaac.core.utils.torch"].maths = maths
    ts = types.ModuleType("torch_scatter")
    ts.scatter = _scatter
    sys.modules.setdefault("torch_scatter", ts)
    for p in (_ENV_DIR, _RL_DIR, os.path.join(_ENV_DIR, "scripts")):
        if p not in sys.path:
            sys.path.insert(0, p)
    _INSTALLED = True


# ----------------------------------------------------------------------------- fake PhysX views
def _dof_index(name: str) -> int:
    """DOF order consistent with the tasks' hard-coded indices (QPC:466-469): dof1 of a1..a4 -> 0..3,
    a_k dof2 -> 4+2(k-1), a_k dof3 -> 5+2(k-1); 12..19 passive."""
    k = int(name[1])
    j = int(name[-1])
    return (k - 1) if j == 1 else (4 + 2 * (k - 1) + (j - 2))


class FakeArticulationView:
    num_dof = 20

    def __init__(self, state: Dict[str, torch.Tensor]):
        self.s = state
        self.calls: List[tuple] = []
        n = state["dof_pos"].shape[0]
        dev = state["dof_pos"].device
        self.targets = {k: torch.zeros(n, 20, device=dev) for k in ("position", "velocity", "effort")}

    def get_dof_index(self, name):
        return _dof_index(name)

    def get_dof_limits(self):
        n = self.s["dof_pos"].shape[0]
        lim = torch.empty(n, 20, 2)
        lim[..., 0], lim[..., 1] = -np.pi, np.pi
        return lim

    def _sel(self, t, joint_indices):
        return t if joint_indices is None else t[:, joint_indices]

    def get_joint_positions(self, indices=None, joint_indices=None, clone=True):
        return self._sel(self.s["dof_pos"], joint_indices)

    def get_joint_velocities(self, indices=None, joint_indices=None, clone=True):
        return self._sel(self.s["dof_vel"], joint_indices)

    def get_world_poses(self, indices=None, clone=True):
        return self.s["root_pos"], self.s["root_quat"]

    def get_linear_velocities(self, indices=None, clone=True):
        return self.s["root_lin_vel"]

    def get_angular_velocities(self, indices=None, clone=True):
        return self.s["root_ang_vel"]

    def _write(self, key, values, indices, joint_indices):
        dst = self.s[key]
        rows = torch.arange(dst.shape[0], device=dst.device) if indices is None else indices
        if joint_indices is None:
            dst[rows] = values
        else:
            dst[rows[:, None], joint_indices[None, :]] = values

    def set_joint_positions(self, positions, indices=None, joint_indices=None):
        self.calls.append(("set_joint_positions", None if indices is None else indices.clone()))
        self._write("dof_pos", positions, indices, joint_indices)

    def set_joint_velocities(self, velocities, indices=None, joint_indices=None):
        self.calls.append(("set_joint_velocities", None if indices is None else indices.clone()))
        self._write("dof_vel", velocities, indices, joint_indices)

    def set_world_poses(self, positions=None, orientations=None, indices=None):
        self.calls.append(("set_world_poses", None if indices is None else indices.clone()))
        rows = slice(None) if indices is None else indices
        if positions is not None:
            self.s["root_pos"][rows] = positions
        if orientations is not None:
            self.s["root_quat"][rows] = orientations

    def set_velocities(self, velocities, indices=None):
        self.calls.append(("set_velocities", None if indices is None else indices.clone()))
        rows = slice(None) if indices is None else indices
        self.s["root_lin_vel"][rows] = velocities[:, 0:3]
        self.s["root_ang_vel"][rows] = velocities[:, 3:6]

    def _target(self, mode, values, joint_indices):
        if joint_indices is None:
            self.targets[mode][:] = values
        else:
            self.targets[mode][:, joint_indices] = values

    def set_joint_position_targets(self, positions, indices=None, joint_indices=None):
        self._target("position", positions, joint_indices)

    def set_joint_velocity_targets(self, velocities, indices=None, joint_indices=None):
        self._target("velocity", velocities, joint_indices)

    def set_joint_efforts(self, efforts, indices=None, joint_indices=None):
        self._target("effort", efforts, joint_indices)

    def set_gains(self, *a, **k):
        pass

    def set_max_efforts(self, *a, **k):
        pass

    def set_friction_coefficients(self, *a, **k):
        pass

    def switch_control_mode(self, *a, **k):
        pass


class FakeRigidPrimView:
    def __init__(self, pos, quat=None, lin=None, ang=None):
        self.pos, self.quat, self.lin, self.ang = pos, quat, lin, ang
        self.calls: List[tuple] = []

    def get_world_poses(self, indices=None, clone=True):
        return self.pos, self.quat

    def get_linear_velocities(self, indices=None, clone=True):
        return self.lin if self.lin is not None else torch.zeros_like(self.pos)

    def get_angular_velocities(self, indices=None, clone=True):
        return self.ang if self.ang is not None else torch.zeros_like(self.pos)

    def set_world_poses(self, positions=None, orientations=None, indices=None):
        self.calls.append(("set_world_poses", None if indices is None else indices.clone()))
        rows = slice(None) if indices is None else indices
        if positions is not None:
            self.pos[rows] = positions
        if orientations is not None:
            self.quat[rows] = orientations

    def set_velocities(self, velocities, indices=None):
        self.calls.append(("set_velocities", None if indices is None else indices.clone()))
        rows = slice(None) if indices is None else indices
        self.lin[rows] = velocities[:, 0:3]
        self.ang[rows] = velocities[:, 3:6]


class _World:
    def is_playing(self):
        return True


class _Env:
    _render = False

    def __init__(self):
        self._world = _World()


class _SimConfig:
    def __init__(self, name, cfg, num_envs, device):
        self.config = {"test": False, "sim_device": device, "rl_device": device, "seed": 42, "task_name": name,
                       "task": {"env": {"clipObservations": cfg.clip_observations, "clipActions": cfg.clip_actions,
                                        "controlFrequencyInv": cfg.control_frequency_inv}}}
        self.task_config = {"env": {"numEnvs": num_envs, "envSpacing": cfg.env_spacing},
                            "sim": {"dt": cfg.dt, "max_episode_length": cfg.max_episode_length}}


_TASK_CLASSES = {
    "QuadrupedPoseControl": ("tasks.quadruped_pose_control_tasks.quadruped_pose_control", "QuadrupedPoseControl"),
    "QuadrupedPoseControlVertical": ("tasks.quadruped_pose_control_tasks.quadruped_pose_control_vertical", "QuadrupedPoseControlVertical"),
    "QuadrupedManipulatePlate": ("tasks.quadruped_manipulate_plate.quadruped_manipulate_plate", "QuadrupedManipulatePlate"),
    "QuadrupedManipulatePlateVertical": ("tasks.quadruped_manipulate_plate.quadruped_manipulate_plate_vertical", "QuadrupedManipulatePlateVertical"),
    "JointLocomanipulation": ("tasks.joint_train_locomanipulation.joint_locomanipulation", "JointLocomanipulation"),
    "JointLocomanipulationVertical": ("tasks.joint_train_locomanipulation.joint_locomanipulation_vertical", "JointLocomanipulationVertical"),
    "QuadrupedPoseControlPositionControl": ("tasks.quadruped_pose_control_tasks.quadruped_pose_control_position_control", "QuadrupedPoseControlPositionControl"),
    "QuadrupedManipulatePlatePositionControl": ("tasks.quadruped_manipulate_plate.quadruped_manipulate_plate_position_control", "QuadrupedManipulatePlatePositionControl"),
    "JointLocomanipulationPositionControl": ("tasks.joint_train_locomanipulation.joint_locomanipulation_position_control", "JointLocomanipulationPositionControl"),
    "QuadrupedPoseControlCustomController": ("tasks.quadruped_pose_control_tasks.quadruped_pose_control_custom_controller", "QuadrupedPoseControlCustomController"),
    "QuadrupedManipulatePlateCustomController": ("tasks.quadruped_manipulate_plate.quadruped_manipulate_plate_custom_controller", "QuadrupedManipulatePlateCustomController"),
}


def make_reference_task(cfg, scene, device: str = "cpu"):
    """Instantiate the reference class for ``cfg.name`` over ``scene``'s tensors."""
    import importlib
    install()
    mod_name, cls_name = _TASK_CLASSES[cfg.name]
    with open(os.devnull, "w") as devnull, contextlib.redirect_stdout(devnull):
        mod = importlib.import_module(mod_name)
        cls = getattr(mod, cls_name)
        if hasattr(cls, "RECORD_JOINT"):
            cls.RECORD_JOINT = False  # JL:37 — else np.save into /home/bionicdl/... and exit(0) (JL:861-874)
        task = cls(sim_config=_SimConfig(cfg.name, cfg, scene.num_envs, device), name=cfg.name, env=_Env())
        task._env_pos = scene.env_pos
        if scene.loco:
            r = task.robot_locomotion
            r._robot_articulation = FakeArticulationView(scene.loco)
            r._tip_view = FakeRigidPrimView(scene.loco["tip_pos"])
            r._knee_view = FakeRigidPrimView(scene.loco["knee_pos"])
            task.pose_indicator_loco.object_view = FakeRigidPrimView(scene.loco["indicator_pos"], scene.loco["indicator_quat"])
        if scene.mani:
            r = task.robot_manipulation
            r._robot_articulation = FakeArticulationView(scene.mani)
            r._tip_view = FakeRigidPrimView(scene.mani["tip_pos"])
            r._knee_view = FakeRigidPrimView(scene.mani["knee_pos"])
            task.obj.object_view = FakeRigidPrimView(scene.mani["plate_pos"], scene.mani["plate_quat"],
                                                     scene.mani["plate_lin_vel"], scene.mani["plate_ang_vel"])
            task.pose_indicator_mani.object_view = FakeRigidPrimView(scene.mani["indicator_pos"], scene.mani["indicator_quat"])
        task.post_reset()
    return task


@contextlib.contextmanager
def supplied_uniforms(buffer: torch.Tensor):
    """Make ``torch.rand((k, 3))`` inside ``rand_quaternions`` (utils/math.py:184) return successive rows of
    ``buffer`` — the kernels receive the same rows and index them by reset rank."""
    real = torch.rand
    cursor = {"i": 0}

    def fake(*size, **kw):
        shape = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else size
        k = int(shape[0])
        i = cursor["i"]
        cursor["i"] = i + k
        return buffer[i:i + k].clone().to(kw.get("device", "cpu"))

    torch.rand = fake
    try:
        yield cursor
    finally:
        torch.rand = real


@contextlib.contextmanager
def quiet():
    with open(os.devnull, "w") as devnull, contextlib.redirect_stdout(devnull):
        yield
