"""Host-side mirror of the reference task surface (tasks/base/rl_task.py:45-260, joint_rl_task.py).

``LocoManipulationTask`` keeps the OmniIsaacGymEnvs task API — ``pre_physics_step(actions)``,
``post_physics_step() -> (obs_buf, rew_buf, reset_buf, extras)``, ``get_observations``, ``get_states``,
``calculate_metrics``, ``is_done``, ``get_extras``, ``reset``, ``reset_idx``, ``post_reset``, the
``obs_buf/states_buf/rew_buf/reset_buf/progress_buf/extras`` buffers with the reference dtypes, and the
attributes the vec-env adapter reads (``clip_obs``, ``clip_actions``, ``rl_device``, ``device``,
``control_frequency_inv``, ``randomize_actions/observations``) — but every method is a thin call into
liblm_b200.so (include/lm_b200.h).  PyTorch owns the tensors; there is no eager fallback: without a
CUDA device or without the built library construction raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from .. import _lib
from ..configs import (CONTROL_EFFORT, CONTROL_PD, KIND_COTRAIN, KIND_LOCO, KIND_MANI, TaskConfig, get_task_config)

_ART_KEYS = ("dof_pos", "dof_vel", "root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "tip_pos", "knee_pos",
             "indicator_pos", "indicator_quat", "plate_pos", "plate_quat", "plate_lin_vel", "plate_ang_vel")


class LocoManipulationTask:
    """One task instance over ``num_envs`` environments on one GPU."""

    def __init__(self, cfg: TaskConfig, num_envs: int, scene=None, device="cuda:0", rl_device=None,
                 clamp_actions: bool = True, fuse_obs_clamp: bool = False, seed: int = 42):
        self.cfg = cfg
        self._name = cfg.name
        self._num_envs = int(num_envs)
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise _lib.LmError("LocoManipulationTask runs on a CUDA device only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise _lib.LmError("CUDA is not available; the B200 task kernels cannot run")
        self._lib = _lib.load()
        self.rl_device = rl_device or str(self._device)
        self._num_actions, self._num_observations, self._num_states = cfg.num_actions, cfg.num_observations, cfg.num_states
        self._num_agents = 1
        self.clip_obs, self.clip_actions = cfg.clip_observations, cfg.clip_actions
        self.control_frequency_inv = cfg.control_frequency_inv
        self.randomize_actions = False        # rl_task.py:72-73 (domain randomisation is off in every hot-path yaml)
        self.randomize_observations = False
        self._dr_randomizer = None
        self._max_episode_length = cfg.max_episode_length
        self._single_task_num_envs = self._num_envs // 2 if cfg.kind == KIND_COTRAIN else self._num_envs
        self._generator = torch.Generator(device=self._device)
        self._generator.manual_seed(seed)
        idx = self._device.index if self._device.index is not None else torch.cuda.current_device()
        # *_custom_controller files: the shared kernels leave the action-rate term to lm_cc_post (|a| instead of |a_last - a|)
        kcfg = cfg.with_(action_rate_scale=0.0) if cfg.custom_controller else cfg
        if cfg.custom_controller and (cfg.kind == KIND_COTRAIN or cfg.num_states != 93 or cfg.quaternion_obs_scale != 1.0):
            raise _lib.LmError("custom_controller: single-task files with the 93-wide states row and quaternion scale 1 only")
        if cfg.custom_controller and cfg.clip_reward and cfg.fall_penalty != 0.0:
            raise _lib.LmError("clip_reward with a non-zero fall_penalty is not supported (the reference clips before adding it)")
        self._fuse_obs_clamp = bool(fuse_obs_clamp)
        self._params = _lib.task_params(kcfg, self._num_envs, clamp_actions=clamp_actions)
        with torch.cuda.device(idx):
            self._handle = self._lib.lm_task_create(C.byref(self._params), idx)
        if not self._handle:
            raise _lib.LmError(f"lm_task_create failed: {self._lib.lm_last_error().decode()}")
        if fuse_obs_clamp:
            _lib.check(self._lib.lm_task_set_clamp_observations(self._handle, 1), "lm_task_set_clamp_observations")
        self.cleanup()
        self._scene = None
        self._loco_views = self._mani_views = None
        self._posted_since_pre = False
        self._reset_version = -1
        if scene is not None:
            self.set_up_scene(scene)

    # ------------------------------------------------------------------ buffers (rl_task.py:104-113; QPC:123-154)
    def cleanup(self) -> None:
        n, dev, f32, i64 = self._num_envs, self._device, torch.float32, torch.long
        self.obs_buf = torch.zeros((n, self._num_observations), device=dev, dtype=f32)
        if self._num_states == 93:
            self.states_buf = torch.zeros((n, 93), device=dev, dtype=f32)
        else:
            self.states_buf = self.obs_buf                      # JLV:515
        self.rew_buf = torch.zeros(n, device=dev, dtype=f32)
        self.reset_buf = torch.ones(n, device=dev, dtype=i64)
        self.progress_buf = torch.zeros(n, device=dev, dtype=i64)
        self.successes = torch.zeros(n, device=dev, dtype=i64)
        self.consecutive_successes = torch.zeros(n, device=dev, dtype=i64)
        self.goal_reset_buf = torch.zeros(n, device=dev, dtype=i64)
        self._actions_ring = torch.zeros((2, n, 12), device=dev, dtype=f32)
        self.last_joint_velocities = torch.zeros((n, 12), device=dev, dtype=f32)
        self.last_base_tip_positions = torch.zeros((n, 4, 3), device=dev, dtype=f32)
        self.goal_quaternions = torch.zeros((n, 4), device=dev, dtype=f32)
        self.reset_env_ids_buf = torch.zeros(n, device=dev, dtype=i64)
        self.num_reset_buf = torch.zeros(2, device=dev, dtype=torch.int32)
        self._extras_buf = torch.zeros(_lib.NUM_EXTRAS, device=dev, dtype=f32)
        self._counters = torch.zeros(6, device=dev, dtype=i64)
        h = self._single_task_num_envs
        tw = 20 if self.cfg.control_mode == CONTROL_EFFORT else 12
        self.targets_loco = torch.zeros((h, tw), device=dev, dtype=f32)
        self.targets_mani = torch.zeros((h, tw), device=dev, dtype=f32)
        ex = self._extras_buf
        self._pd = None
        if self.cfg.control_mode == CONTROL_PD:
            self._setup_pd_state()
        self.extras: Dict[str, torch.Tensor] = {"env/rewards/" + k: ex[i] for i, k in enumerate(_lib.EXTRA_NAMES)}
        self._cc = None
        if self.cfg.custom_controller:                                   # …custom_controller.py:161,188,610-612
            self._obs64 = torch.zeros((n, 64), device=dev, dtype=f32)    # what the shared k_post writes; obs_buf is the 88-wide row
            self.extras["env/rewards/mechanical_power_penalty"] = ex[12]
            self.extras["env/rewards/position_target_error_penalty"] = ex[13]
            if self.cfg.cc_rot_dist_term:
                self.last_rot_dist = torch.zeros(n, device=dev, dtype=f32)
                self.extras["env/rewards/rot_dist_decreasing_reward"] = ex[14]
        self.extras["env/success_rate"] = ex[_lib.EXTRA_SUCCESS_RATE]
        if self.cfg.kind == KIND_COTRAIN:
            self.extras["env/success_rate_loco"] = ex[_lib.EXTRA_SUCCESS_RATE + 1]
            self.extras["env/success_rate_mani"] = ex[_lib.EXTRA_SUCCESS_RATE + 2]
        b = _lib.LmTaskBuffers()
        b.obs_buf, b.rew_buf = (self._obs64 if self.cfg.custom_controller else self.obs_buf).data_ptr(), self.rew_buf.data_ptr()
        b.states_buf = self.states_buf.data_ptr() if self._num_states == 93 else None
        b.reset_buf, b.progress_buf = self.reset_buf.data_ptr(), self.progress_buf.data_ptr()
        b.successes, b.consecutive_successes = self.successes.data_ptr(), self.consecutive_successes.data_ptr()
        b.goal_reset_buf, b.actions_ring = self.goal_reset_buf.data_ptr(), self._actions_ring.data_ptr()
        b.last_joint_vel, b.last_base_tip = self.last_joint_velocities.data_ptr(), self.last_base_tip_positions.data_ptr()
        b.goal_quaternions, b.reset_env_ids = self.goal_quaternions.data_ptr(), self.reset_env_ids_buf.data_ptr()
        b.num_reset, b.extras, b.counters = self.num_reset_buf.data_ptr(), self._extras_buf.data_ptr(), self._counters.data_ptr()
        self._buffers = b

    # ------------------------------------------------------------------ *_position_control state
    def _setup_pd_state(self) -> None:
        """Buffers of the explicit-PD files (quadruped_pose_control_position_control.py:179-191;
        joint_locomanipulation_position_control.py:161-183), env-indexed [N,12] with locomotion rows first."""
        cfg, n, dev, f32 = self.cfg, self._num_envs, self._device, torch.float32
        init = torch.zeros((n, 12), device=dev, dtype=f32)
        h = self._single_task_num_envs
        loco = torch.tensor(cfg.loco_init_joint_pos[:12], device=dev, dtype=f32)
        mani = torch.tensor(cfg.mani_init_joint_pos[:12], device=dev, dtype=f32)
        if cfg.kind == KIND_COTRAIN:
            init[:h], init[h:] = loco, mani
        else:
            init[:] = loco if cfg.kind == KIND_LOCO else mani
        self.joint_positions = init.clone()                       # joint_positions_loco / _mani / _combined
        self.joint_velocities = torch.zeros((n, 12), device=dev, dtype=f32)
        self.current_joint_position_targets = init.clone()
        self.last_joint_position_targets = init.clone()
        self.current_joint_position_targets_se = torch.tensor(cfg.pd_init_se, device=dev, dtype=f32).repeat(n, 1)
        self.applied_torque = torch.zeros((n, 12), device=dev, dtype=f32)   # last set_joint_efforts value
        self._torque_sum = torch.zeros((n, 12), device=dev, dtype=f32)
        self.torque = torch.zeros((n, 12), device=dev, dtype=f32)           # …position_control.py:313
        self._se_lower = (C.c_float * 12)(*cfg.pd_min_se)
        self._se_upper = (C.c_float * 12)(*cfg.pd_max_se)
        pd = _lib.LmPdState()
        pd.joint_pos, pd.joint_vel = self.joint_positions.data_ptr(), self.joint_velocities.data_ptr()
        pd.targets, pd.last_targets = self.current_joint_position_targets.data_ptr(), self.last_joint_position_targets.data_ptr()
        pd.targets_se = self.current_joint_position_targets_se.data_ptr()
        for j in range(12):
            pd.init_se[j] = cfg.pd_init_se[j]
        pd.target_obs_scale = cfg.pd_target_obs_scale
        pd.carry_post_state = int(cfg.pd_carry_post_state)
        _lib.check(self._lib.lm_task_set_pd_state(self._handle, C.byref(pd)), "lm_task_set_pd_state")
        self._pd = pd

    def _cc_struct(self):
        """LmCustomController over the current scene views (rebuilt when the scene changes)."""
        cfg = self.cfg
        st = self._scene.loco if cfg.kind == KIND_LOCO else self._scene.mani
        if self._cc is not None and self._cc_scene is self._scene:
            return self._cc
        c = _lib.LmCustomController()
        c.num_envs, c.clip_reward, c.clamp_observations = self._num_envs, int(cfg.clip_reward), int(self._fuse_obs_clamp)
        c.flags = (1 if cfg.cc_update_last_targets else 0) | (2 if cfg.cc_rot_dist_term else 0)
        c.clip_observations, c.target_obs_scale = cfg.clip_observations, cfg.pd_target_obs_scale
        c.action_rate_scale, c.mechanical_power_penalty_scale = cfg.action_rate_scale, cfg.mechanical_power_penalty_scale
        c.position_target_error_penalty_scale = cfg.position_target_error_penalty_scale
        c.rot_dist_decreasing_reward_scale = cfg.rot_dist_decreasing_reward_scale
        c.no_rot_dist_decreasing_reward_thresh = cfg.no_rot_dist_decreasing_reward_thresh
        for i in range(4):
            c.default_quaternion[i] = cfg.loco_default_quaternion[i]          # lm_cc_pre (locomotion file only)
        for j in range(12):
            c.dof_index[j] = self._params.dof_index[j]
        c.obs64 = self._obs64.data_ptr()
        c.dof_pos, c.dof_vel, c.goal = st["dof_pos"].data_ptr(), st["dof_vel"].data_ptr(), self.goal_quaternions.data_ptr()
        c.torque, c.targets, c.last_targets = self.torque.data_ptr(), self.current_joint_position_targets.data_ptr(), self.last_joint_position_targets.data_ptr()
        c.joint_pos, c.joint_vel = self.joint_positions.data_ptr(), self.joint_velocities.data_ptr()
        c.last_rot_dist = self.last_rot_dist.data_ptr() if cfg.cc_rot_dist_term else None
        c.obs, c.states, c.rew, c.extras = self.obs_buf.data_ptr(), self.states_buf.data_ptr(), self.rew_buf.data_ptr(), self._extras_buf.data_ptr()
        c.reset_env_ids, c.num_reset = self.reset_env_ids_buf.data_ptr(), self.num_reset_buf.data_ptr()
        if getattr(self, "_cc_ws", None) is None:
            from ..rollout import _handle
            _, self._cc_ws = _handle(self._device)
        self._cc, self._cc_scene = c, self._scene
        return c

    def _pd_groups(self):
        """(articulation state dict, first env row, rows) per robot."""
        h = self._single_task_num_envs
        out = []
        if self.cfg.kind != KIND_MANI:
            out.append((self._scene.loco, 0, h))
        if self.cfg.kind != KIND_LOCO:
            out.append((self._scene.mani, h if self.cfg.kind == KIND_COTRAIN else 0, h))
        return out

    def _pd_loop(self, actions: torch.Tensor, physics_substep) -> None:
        """…position_control.py:261-313: integrate the swing/extension targets, then control_decimal x (torque, physics
        sub-step, joint-state refresh) and one final torque.  ``physics_substep(i)`` stands in for
        ``SimulationContext.step`` (PhysX is out of scope): it must leave the articulation tensors holding the new state."""
        cfg, lib, s = self.cfg, self._lib, self._stream()
        n = self._num_envs
        _lib.check(lib.lm_pd_integrate_targets(actions.data_ptr(), self.current_joint_position_targets_se.data_ptr(),
                                               self.current_joint_position_targets.data_ptr(), self._se_lower, self._se_upper, n,
                                               cfg.pd_action_scale, int(cfg.zero_actions), s), "lm_pd_integrate_targets")
        self._torque_sum.zero_()
        _lib.check(lib.lm_pd_torque(self.joint_positions.data_ptr(), self.joint_velocities.data_ptr(),
                                    self.current_joint_position_targets.data_ptr(), self.applied_torque.data_ptr(),
                                    self._torque_sum.data_ptr(), n, cfg.pd_kp, cfg.pd_kd, cfg.pd_max_effort, s), "lm_pd_torque")
        for i in range(cfg.pd_control_decimal):
            if physics_substep is not None:
                physics_substep(i)
            last = i == cfg.pd_control_decimal - 1
            for st, row0, rows in self._pd_groups():
                sl = slice(row0, row0 + rows)
                rc = lib.lm_pd_update_and_torque(st["dof_pos"].data_ptr(), st["dof_vel"].data_ptr(), None,
                                                 self.joint_positions[sl].data_ptr(), self.joint_velocities[sl].data_ptr(),
                                                 self.last_joint_velocities[sl].data_ptr(), self.current_joint_position_targets[sl].data_ptr(),
                                                 self.applied_torque[sl].data_ptr(), self._torque_sum[sl].data_ptr(),
                                                 int(cfg.pd_sum_final_torque) if last else 1,
                                                 self.torque[sl].data_ptr() if last else None, cfg.pd_torque_divisor, rows,
                                                 cfg.pd_kp, cfg.pd_kd, cfg.pd_max_effort, s)
                _lib.check(rc, "lm_pd_update_and_torque")

    # ------------------------------------------------------------------ scene / views
    def _views(self, state: Optional[Dict[str, torch.Tensor]], targets: torch.Tensor):
        if not state:
            return None
        v = _lib.LmRobotViews()
        for k in _ART_KEYS:
            t = state.get(k)
            if t is not None:
                if t.device != self._device or t.dtype != torch.float32 or not t.is_contiguous():
                    raise _lib.LmError(f"view tensor {k} must be a contiguous float32 tensor on {self._device}")
                setattr(v, k, t.data_ptr())
        v.targets = targets.data_ptr()
        return v

    def set_up_scene(self, scene) -> None:
        """Bind the PhysX-side tensors (``scene.loco`` / ``scene.mani`` dicts and ``scene.env_pos``).  The reference's
        set_up_scene clones USD prims and creates the views (QPC:161-186); only the resulting tensors matter here."""
        self._scene = scene
        self._env_pos = scene.env_pos
        assert self._env_pos.shape[0] == self._single_task_num_envs
        self._loco_views = self._views(scene.loco, self.targets_loco)
        self._mani_views = self._views(scene.mani, self.targets_mani)
        if self.cfg.kind != KIND_MANI and self._loco_views is None:
            raise _lib.LmError("scene has no locomotion robot state")
        if self.cfg.kind != KIND_LOCO and self._mani_views is None:
            raise _lib.LmError("scene has no manipulation robot state")

    def post_reset(self) -> None:
        """QPC:188-198 — limits / gains / DOF index discovery are PhysX plumbing; nothing to compute."""
        return None

    # ------------------------------------------------------------------ properties (rl_task.py:166-219)
    num_envs = property(lambda self: self._num_envs)
    num_actions = property(lambda self: self._num_actions)
    num_observations = property(lambda self: self._num_observations)
    num_states = property(lambda self: self._num_states)
    num_agents = property(lambda self: self._num_agents)
    device = property(lambda self: str(self._device))

    @property
    def current_actions(self) -> torch.Tensor:
        return self._actions_ring[self._lib.lm_task_step_parity(self._handle)]

    @property
    def last_actions(self) -> torch.Tensor:
        p = self._lib.lm_task_step_parity(self._handle)
        # after post_physics_step the reference has last_actions == current_actions (QPC:548)
        return self._actions_ring[p if self._posted_since_pre else p ^ 1]

    @property
    def num_successes(self):
        return self._counters[0]

    @property
    def num_resets(self):
        return self._counters[1]

    def _stream(self):
        return torch.cuda.current_stream(self._device).cuda_stream

    def _lp(self):
        return C.byref(self._loco_views) if self._loco_views is not None else None

    def _mp(self):
        return C.byref(self._mani_views) if self._mani_views is not None else None

    # ------------------------------------------------------------------ the task API
    def get_states(self):
        return self.states_buf

    def get_extras(self):
        return self.extras

    def enable_noise_randomization(self, randomization_params: dict, seed: int = 0) -> None:
        """Observation / action noise of the vec-env step: ``randomization_params`` is the ``observations`` / ``actions`` part of
        the task yaml's ``domain_randomization.randomization_params`` (cfg/task/JointLocomanipulation.yaml:171-191).  Sets the
        attributes ``VecEnvRLGames.step`` reads (envs/vec_env_rlgames.py:57-58,70-72; randomize.py:156-211)."""
        from ..noise import Randomizer
        self._dr_randomizer = Randomizer(randomization_params, self._num_envs, self.num_observations, self.num_actions, device=self._device, seed=seed)
        self.randomize_observations = self._dr_randomizer.randomize_observations
        self.randomize_actions = self._dr_randomizer.randomize_actions

    def reset(self) -> None:
        """rl_task.py:222-225 (in place, so the pointer the kernels hold stays valid)."""
        self.reset_buf.fill_(1)

    def draw_goal_uniforms(self) -> torch.Tensor:
        """The ``torch.rand((num_reset, 3))`` draw of rand_quaternions (utils/math.py:184), drawn for every env up
        front so no host sync on ``num_reset`` is needed; row k is consumed by the k-th reset env."""
        return torch.rand((self._num_envs, 3), device=self._device, generator=self._generator)

    def action_sink(self) -> "_lib.LmActionSink":
        """The action half of the NEXT ``pre_physics_step`` as a struct the producing kernel takes (``MlpActorCritic.forward(sink=)``,
        ``RolloutBuffer.write(sink=)``): clamp, ``current_actions`` stash and ``take_action`` then happen inside that kernel, and the
        next ``pre_physics_step(None, uniforms)`` runs only its reset half.  Call after ``post_physics_step`` (it depends on the
        ring parity and on the scene bound at that moment)."""
        sink = _lib.LmActionSink()
        _lib.check(self._lib.lm_task_action_sink(self._handle, self._lp(), self._mp(), C.byref(self._buffers), C.byref(sink)), "lm_task_action_sink")
        return sink

    def pre_physics_step(self, actions: Optional[torch.Tensor], uniforms: Optional[torch.Tensor] = None, physics_substep=None) -> None:
        """``actions is None``: they were delivered through :meth:`action_sink` by the kernel that produced them."""
        if self._scene is None:
            raise _lib.LmError("set_up_scene(scene) must be called before stepping")
        if actions is not None:
            actions = actions.to(self._device, dtype=torch.float32)
            if not actions.is_contiguous():
                actions = actions.contiguous()
        if uniforms is None:
            uniforms = self.draw_goal_uniforms()
        if self.reset_buf._version != self._reset_version:
            # reset_buf was written through torch since the last post step (task.reset(), user code): rescan it
            self._lib.lm_task_invalidate_reset_scan(self._handle)
        with _lib.nvtx("pre_physics_step"):
            rc = self._lib.lm_task_pre_step(self._handle, self._lp(), self._mp(), self._env_pos.data_ptr(), C.byref(self._buffers),
                                            _lib.ptr(actions), uniforms.data_ptr(), self._stream())
        _lib.check(rc, "lm_task_pre_step")
        if self.cfg.custom_controller and self.cfg.cc_rot_dist_term:     # last_rot_dist of the envs reset just now (…custom_controller.py:384-386)
            _lib.check(self._lib.lm_cc_pre(C.byref(self._cc_struct()), self._stream()), "lm_cc_pre")
        if self._pd is not None:
            self._pd_loop(self.current_actions, physics_substep)   # the stashed (adapter-clamped) actions
        self._posted_since_pre = False
        self._keep = (actions, uniforms)   # keep inputs alive until the stream has consumed them

    def reset_idx(self, env_ids: torch.Tensor, uniforms: Optional[torch.Tensor] = None) -> None:
        env_ids = env_ids.to(self._device, dtype=torch.long).contiguous()
        k = int(env_ids.numel())
        if k == 0:
            return
        if uniforms is None:
            uniforms = torch.rand((k, 3), device=self._device, generator=self._generator)
        rc = self._lib.lm_task_reset_idx(self._handle, self._lp(), self._mp(), self._env_pos.data_ptr(), C.byref(self._buffers),
                                         env_ids.data_ptr(), k, uniforms.data_ptr(), self._stream())
        _lib.check(rc, "lm_task_reset_idx")
        self._keep = (env_ids, uniforms)

    def set_defer_finish(self, on: bool = True) -> None:
        """``on``: ``post_physics_step`` only publishes per-tile partial sums and reset counts; extras, success counters and
        ``num_reset`` are produced by one extra CTA inside the NEXT ``pre_physics_step``'s launch (beside the reset bodies) instead of
        as the serial tail of the post kernel, or by :meth:`post_finish` when they are wanted earlier (rollout boundaries, logging)."""
        _lib.check(self._lib.lm_task_set_defer_finish(self._handle, int(bool(on))), "lm_task_set_defer_finish")
        self._defer_finish = bool(on)

    def set_early_reset(self, on: bool = True) -> None:
        """With an action sink: ``pre_physics_step(None, uniforms)`` (the reset half) runs beside the kernel launched before it — the
        policy forward delivering the actions — instead of behind it (see ``lm_task_set_early_reset`` for the contract)."""
        _lib.check(self._lib.lm_task_set_early_reset(self._handle, int(bool(on))), "lm_task_set_early_reset")

    def post_finish(self) -> None:
        """Run the deferred finish pass now (acts at most once per post step)."""
        _lib.check(self._lib.lm_task_post_finish(self._handle, C.byref(self._buffers), self._stream()), "lm_task_post_finish")

    def post_physics_step(self):
        """rl_task.py:240-260: progress += 1, observations, metrics, termination — one fused launch."""
        if self._posted_since_pre:
            # second post without a pre: the reference would see last_actions == current_actions (QPC:548)
            p = self._lib.lm_task_step_parity(self._handle)
            self._actions_ring[p ^ 1].copy_(self._actions_ring[p])
        with _lib.nvtx("post_physics_step"):
            rc = self._lib.lm_task_post_step(self._handle, self._lp(), self._mp(), self._env_pos.data_ptr(), C.byref(self._buffers),
                                             self._stream())
        _lib.check(rc, "lm_task_post_step")
        if self.cfg.custom_controller:
            c = self._cc_struct()
            _lib.check(self._lib.lm_cc_post(self._cc_ws, C.byref(c), self._stream()), "lm_cc_post")
        self._posted_since_pre = True
        self._reset_version = self.reset_buf._version
        return self.obs_buf, self.rew_buf, self.reset_buf, self.extras

    # The reference runs these three in sequence inside post_physics_step; here they are one kernel.  Calling
    # get_observations() runs it; the other two are then no-ops on already-computed buffers.
    def get_observations(self):
        self.progress_buf -= 1      # post_physics_step increments; a bare get_observations() must not (QPC:301)
        self.post_physics_step()
        return self.obs_buf

    def calculate_metrics(self) -> None:
        return None

    def is_done(self) -> None:
        return None

    def reset_env_ids(self) -> torch.Tensor:
        """Ascending ids reset by the last pre_physics_step (host sync; for logging / parity checks).  Valid between
        pre_physics_step and post_physics_step: the post step overwrites the count with the NEXT step's reset count."""
        k = int(self.num_reset_buf[0].item())
        return self.reset_env_ids_buf[:k]

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self._lib.lm_task_destroy(self._handle)
                self._handle = None
        except Exception:
            pass


def make_task(name: str, num_envs: int, scene=None, device="cuda:0", **kw) -> LocoManipulationTask:
    """Name -> task, the role of ``initialize_task``'s task_map (utils/task_util.py:60-90)."""
    return LocoManipulationTask(get_task_config(name), num_envs, scene=scene, device=device, **kw)
