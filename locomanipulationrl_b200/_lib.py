"""ctypes binding of liblm_b200.so (include/lm_b200.h).  The structures below mirror the C structs
field by field; tests/test_cabi.py checks sizes and that every declared symbol is exported.

There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from .configs import TaskConfig, DEFAULT_BASE_TIP_POSITIONS, CONTROL_EFFORT, CONTROL_POSITION, CONTROL_VELOCITY

NUM_DOF, NUM_ACT, NUM_OBS, NUM_STATES_FULL, NUM_EXTRAS = 20, 12, 64, 93, 16
_f, _i32, _i64, _fp = C.c_float, C.c_int32, C.c_int64, C.POINTER(C.c_float)
_i64p, _i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)

EXTRA_NAMES = ["orientation_rew", "translation_penalty", "joint_acc_penalty", "action_rate_penalty",
               "consecutive_successes_rew", "joint_limit_panelty", "fall_penalty"]
EXTRA_SUCCESS_RATE, EXTRA_NAN_COUNT, EXTRA_BELOW_M50 = 7, 10, 11


class LmTaskParams(C.Structure):
    _fields_ = [
        ("kind", _i32), ("num_envs", _i32), ("num_states", _i32), ("control_mode", _i32), ("zero_actions", _i32),
        ("joint1_mirrored", _i32), ("max_consecutive_successes", _i32), ("max_episode_length", _i32),
        ("max_reset_counts", _i32), ("clamp_actions", _i32), ("dof_index", _i32 * NUM_ACT),
        ("min_roll", _f), ("roll_range", _f), ("min_pitch", _f), ("pitch_range", _f), ("min_yaw", _f), ("yaw_range", _f),
        ("position_scale", _f), ("quaternion_obs_scale", _f), ("linear_vel_scale", _f), ("angular_vel_scale", _f),
        ("tip_position_scale", _f), ("joint_position_scale", _f), ("joint_velocity_scale", _f),
        ("quaternion_scale", _f), ("rot_eps", _f), ("translation_scale", _f), ("joint_acc_scale", _f),
        ("action_rate_scale", _f), ("consecutive_success_reward", _f), ("fall_penalty", _f), ("joint_limit_penalty", _f),
        ("baseline_knee_height", _f), ("baseline_height", _f), ("baseline_corner_height", _f), ("success_thresh", _f),
        ("min_joint_23_diff", _f), ("max_joint_23_diff", _f), ("reset_min_joint_23_diff", _f), ("reset_max_joint_23_diff", _f),
        ("min_joint_1_pos", _f), ("max_joint_1_pos", _f), ("reset_min_joint_1_pos", _f), ("reset_max_joint_1_pos", _f),
        ("time_interval", _f), ("clip_actions", _f), ("clip_observations", _f), ("action_limit", _f),
        ("corners", (_f * 3) * 4), ("default_base_tip", (_f * 3) * 4),
        ("loco_init_joint_pos", _f * NUM_DOF), ("loco_default_position", _f * 3), ("loco_default_quaternion", _f * 4),
        ("mani_init_joint_pos", _f * NUM_DOF), ("mani_default_position", _f * 3), ("mani_default_quaternion", _f * 4),
        ("obj_default_position", _f * 3), ("indicator_loco_position", _f * 3), ("indicator_mani_position", _f * 3),
    ]


class LmRobotViews(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "dof_pos", "dof_vel", "root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "tip_pos", "knee_pos",
        "indicator_pos", "indicator_quat", "plate_pos", "plate_quat", "plate_lin_vel", "plate_ang_vel", "targets")]


class LmTaskBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "obs_buf", "states_buf", "rew_buf", "reset_buf", "progress_buf", "successes", "consecutive_successes",
        "goal_reset_buf", "actions_ring", "last_joint_vel", "last_base_tip", "goal_quaternions", "reset_env_ids",
        "num_reset", "extras", "counters")]


class LmHostStep(C.Structure):
    _fields_ = [("actions_host", C.c_void_p), ("uniforms_host", C.c_void_p),
                ("loco_state_host", C.c_void_p), ("loco_state_floats", _i64), ("loco_state_dev", C.c_void_p),
                ("mani_state_host", C.c_void_p), ("mani_state_floats", _i64), ("mani_state_dev", C.c_void_p),
                ("actions_dev", C.c_void_p), ("uniforms_dev", C.c_void_p),
                ("obs_host", C.c_void_p), ("rew_host", C.c_void_p), ("reset_host", C.c_void_p), ("h2d_stream", C.c_void_p)]


class LmHostPolicy(C.Structure):
    _fields_ = [("kind", _i32), ("pad_", _i32), ("handle", C.c_void_p), ("norm_mean", C.c_void_p), ("norm_inv_std", C.c_void_p),
                ("norm_clip", C.c_float), ("pad2_", _i32), ("mu_dev", C.c_void_p), ("value_dev", C.c_void_p),
                ("mu_host", C.c_void_p), ("value_host", C.c_void_p)]


class LmPdState(C.Structure):
    _fields_ = [("joint_pos", C.c_void_p), ("joint_vel", C.c_void_p), ("targets", C.c_void_p), ("last_targets", C.c_void_p),
                ("targets_se", C.c_void_p), ("init_se", C.c_float * NUM_ACT), ("carry_post_state", _i32), ("target_obs_scale", C.c_float)]


class LmRunningStats(C.Structure):
    _fields_ = [("mean", C.c_void_p), ("var", C.c_void_p), ("count", C.c_void_p), ("dim", _i32), ("variant", _i32)]


class LmRolloutStep(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs", "mu", "value", "log_std", "noise", "rew", "dones")] + \
               [("noise_seed", C.c_uint64), ("noise_step", C.c_uint64), ("sample_in_kernel", _i32), ("pad_", _i32),
                ("time_outs", C.c_void_p), ("gamma", _f), ("pad2_", _i32), ("value_denorm", C.c_void_p), ("sink", C.c_void_p)]


class LmRolloutBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obses", "actions", "mus", "values", "rewards", "dones", "neglogp")] + \
               [("horizon", _i32), ("num_envs", _i32), ("reward_scale", _f)]


class LmCustomController(C.Structure):
    _fields_ = [("num_envs", _i32), ("clip_reward", _i32), ("clamp_observations", _i32), ("flags", _i32),
                ("clip_observations", _f), ("target_obs_scale", _f), ("action_rate_scale", _f), ("mechanical_power_penalty_scale", _f),
                ("position_target_error_penalty_scale", _f), ("rot_dist_decreasing_reward_scale", _f),
                ("no_rot_dist_decreasing_reward_thresh", _f), ("default_quaternion", _f * 4), ("dof_index", _i32 * NUM_ACT)] + \
               [(n, C.c_void_p) for n in ("obs64", "dof_pos", "dof_vel", "goal", "torque", "targets", "last_targets", "joint_pos",
                                          "joint_vel", "last_rot_dist", "obs", "states", "rew", "extras", "reset_env_ids", "num_reset")]


class LmNoiseSpec(C.Structure):
    _fields_ = [("enabled", _i32), ("operation", _i32), ("distribution", _i32), ("frequency_interval", _i32), ("scale", _f), ("shift", _f)]


class LmNoiseState(C.Structure):
    _fields_ = [("counter", C.c_void_p), ("correlated", C.c_void_p), ("num_envs", _i32), ("dim", _i32),
                ("on_reset", LmNoiseSpec), ("on_interval", LmNoiseSpec)]


class LmPpoBatch(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("mu", "value", "log_std", "actions", "old_neglogp", "advantages", "old_values", "returns",
                                          "old_mu", "old_log_std")] + \
               [("batch", _i64), ("e_clip", _f), ("critic_coef", _f), ("entropy_coef", _f), ("bounds_loss_coef", _f),
                ("clip_value", _i32), ("update_old", _i32)]


class LmActionSink(C.Structure):
    _fields_ = [("ring_next", C.c_void_p), ("targets_loco", C.c_void_p), ("targets_mani", C.c_void_p), ("loco_envs", _i32),
                ("clamp_actions", _i32), ("zero_actions", _i32), ("pad_", _i32), ("clip_actions", _f), ("action_limit", _f)]


class LmAdamSchedule(C.Structure):
    _fields_ = [("lr", C.c_void_p), ("kl", C.c_void_p), ("kl_scale", _f), ("kl_threshold", _f), ("lr_min", _f), ("lr_max", _f)]


# name -> (restype, argtypes); every symbol include/lm_b200.h declares
_vp = C.c_void_p
SYMBOLS = {
    "lm_abi_version": (C.c_int, []),
    "lm_last_error": (C.c_char_p, []),
    "lm_device_sm_count": (C.c_int, [C.c_int]),
    "lm_struct_size": (C.c_int, [C.c_int]),
    "lm_task_create": (_vp, [C.POINTER(LmTaskParams), C.c_int]),
    "lm_task_destroy": (None, [_vp]),
    "lm_task_step_parity": (C.c_int, [_vp]),
    "lm_task_set_step_parity": (C.c_int, [_vp, C.c_int]),
    "lm_task_set_defer_finish": (C.c_int, [_vp, C.c_int]),
    "lm_task_set_early_reset": (C.c_int, [_vp, C.c_int]),
    "lm_task_post_finish": (C.c_int, [_vp, C.POINTER(LmTaskBuffers), _vp]),
    "lm_task_set_clamp_observations": (C.c_int, [_vp, C.c_int]),
    "lm_task_invalidate_reset_scan": (C.c_int, [_vp]),
    "lm_task_set_timeline": (C.c_int, [_vp, _vp]),
    "lm_task_set_pd_state": (C.c_int, [_vp, C.POINTER(LmPdState)]),
    "lm_pd_torque": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _f, _f, _f, _vp]),
    "lm_pd_update_and_torque": (C.c_int, [_vp, _vp, C.POINTER(_i32), _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _f, _i32, _f, _f, _f, _vp]),
    "lm_task_pre_step": (C.c_int, [_vp, C.POINTER(LmRobotViews), C.POINTER(LmRobotViews), _vp, C.POINTER(LmTaskBuffers), _vp, _vp, _vp]),
    "lm_task_reset_idx": (C.c_int, [_vp, C.POINTER(LmRobotViews), C.POINTER(LmRobotViews), _vp, C.POINTER(LmTaskBuffers), _vp, _i32, _vp, _vp]),
    "lm_task_post_step": (C.c_int, [_vp, C.POINTER(LmRobotViews), C.POINTER(LmRobotViews), _vp, C.POINTER(LmTaskBuffers), _vp]),
    "lm_clamp_copy": (C.c_int, [_vp, _vp, _i64, _f, _vp]),
    "lm_pd_substep": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.POINTER(_i32), _i32, _f, _f, _f, _vp]),
    "lm_pd_integrate_targets": (C.c_int, [_vp, _vp, _vp, C.POINTER(_f), C.POINTER(_f), _i32, _f, _i32, _vp]),
    "lm_mlp_create": (_vp, [C.c_int]),
    "lm_mlp_destroy": (None, [_vp]),
    "lm_mlp_pack": (C.c_int, [_vp] * 11 + [_vp]),
    "lm_mlp_forward": (C.c_int, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _i32, _vp]),
    "lm_mlp_forward_sink": (C.c_int, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _i32, C.POINTER(LmActionSink), _vp]),
    "lm_task_action_sink": (C.c_int, [_vp, C.POINTER(LmRobotViews), C.POINTER(LmRobotViews), C.POINTER(LmTaskBuffers), C.POINTER(LmActionSink)]),
    "lm_mlp_set_timeline": (C.c_int, [_vp, _vp]),
    "lm_gnn_create": (_vp, [C.c_int]),
    "lm_gnn_destroy": (None, [_vp]),
    "lm_gnn_pack": (C.c_int, [_vp, C.POINTER(_vp), _i32, _vp]),
    "lm_gnn_forward": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp]),
    "lm_gnn_forward_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp]),
    "lm_gnn_set_timeline": (C.c_int, [_vp, _vp]),
    "lm_gnn_grad_workspace_floats": (_i32, []),
    "lm_gnn_backward": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _i32, C.POINTER(_vp), _i32, _vp]),
    "lm_rollout_create": (_vp, [C.c_int]),
    "lm_rollout_destroy": (None, [_vp]),
    "lm_rollout_write": (C.c_int, [_vp, C.POINTER(LmRolloutStep), C.POINTER(LmRolloutBuffers), _i32, _vp, C.POINTER(LmRunningStats), _vp]),
    "lm_rollout_finish": (C.c_int, [_vp, C.POINTER(LmRolloutStep), C.POINTER(LmRolloutBuffers), _vp, _vp, _vp]),
    "lm_rms_update": (C.c_int, [_vp, _vp, _i64, C.POINTER(LmRunningStats), _vp]),
    "lm_rms_export_denorm": (C.c_int, [C.POINTER(LmRunningStats), _vp, _vp]),
    "lm_normalize_clamp": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _f, _vp]),
    "lm_moments": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "lm_moment_normalize": (C.c_int, [_vp, _vp, _i64, _vp, _f, _vp]),
    "lm_rms_export": (C.c_int, [C.POINTER(LmRunningStats), _vp, _vp, _vp]),
    "lm_cc_pre": (C.c_int, [C.POINTER(LmCustomController), _vp]),
    "lm_cc_post": (C.c_int, [_vp, C.POINTER(LmCustomController), _vp]),
    "lm_noise_create": (_vp, [C.c_int]),
    "lm_noise_destroy": (None, [_vp]),
    "lm_noise_apply": (C.c_int, [_vp, C.POINTER(LmNoiseState), _vp, _vp, _vp, _vp, _i32, C.c_uint64, C.c_uint64, _vp]),
    "lm_ppo_loss": (C.c_int, [_vp, C.POINTER(LmPpoBatch), _vp, _vp, _vp, _vp, _vp]),
    "lm_mlp_forward_train": (C.c_int, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "lm_mlp_backward": (C.c_int, [_vp] * 10 + [_i32, _vp]),
    "lm_mlp_blob_bytes": (_i64, [_i64, _i32]),
    "lm_mlp_blob_pack": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "lm_mlp_blob_unpack": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "lm_mlp_wgrad_workspace_floats": (_i32, []),
    "lm_mlp_wgrad": (C.c_int, [_vp] * 8 + [_i64, _vp, _i32] + [_vp] * 10 + [_vp]),
    "lm_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i32, _f, _f, _vp, _vp]),
    "lm_adam_step_scheduled": (C.c_int, [_vp, _vp, _vp, _vp, _i64, C.POINTER(LmAdamSchedule), _f, _f, _f, _f, _f, _i32, _f, _f, _vp, _vp]),
    "lm_comm_handle_bytes": (C.c_int, []),
    "lm_comm_create": (_vp, [C.c_int, C.c_int, C.c_int, _i64]),
    "lm_comm_get_handle": (C.c_int, [_vp, _vp]),
    "lm_comm_open": (C.c_int, [_vp, _vp]),
    "lm_comm_destroy": (None, [_vp]),
    "lm_comm_allreduce": (C.c_int, [_vp, _vp, _i64, _vp]),
    "lm_comm_allreduce_gather": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i32, _f, _vp, _i32, _f, _vp]),
    "lm_comm_status": (C.c_int, [_vp]),
    "lm_mlp_wgrad_allreduce": (C.c_int, [_vp] + [_vp] * 8 + [_i64, _vp, _i32] + [_vp] * 10 + [_vp, _i64, _i64, _i64, _i64, _i64, _vp]),
    "lm_gae": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f, _f, _i32, _vp]),
    "lm_gae_env_major": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f, _f, _i32, _vp]),
    "lm_task_step_host": (C.c_int, [_vp, C.POINTER(LmRobotViews), C.POINTER(LmRobotViews), _vp, C.POINTER(LmTaskBuffers), C.POINTER(LmHostStep), _vp]),
    "lm_task_step_host_async": (C.c_int, [_vp, C.POINTER(LmRobotViews), C.POINTER(LmRobotViews), _vp, C.POINTER(LmTaskBuffers), C.POINTER(LmHostStep), _vp]),
    "lm_task_step_host_policy_async": (C.c_int, [_vp, C.POINTER(LmRobotViews), C.POINTER(LmRobotViews), _vp, C.POINTER(LmTaskBuffers), C.POINTER(LmHostStep),
                                                 C.POINTER(LmHostPolicy), _vp]),
    "lm_stream_synchronize": (C.c_int, [_vp]),
    "lm_host_alloc": (_vp, [_i64, _i32]),
    "lm_host_free": (None, [_vp]),
}

# LM_B200_LIB: load another build of the same library (kernel experiments); there is still no fallback
LIB_PATH = os.environ.get("LM_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblm_b200.so")
_lib: Optional[C.CDLL] = None


class LmError(RuntimeError):
    pass


def load(path: Optional[str] = None) -> C.CDLL:
    """Load the C-ABI library; raises if it is missing (no CPU / eager fallback exists)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise LmError(f"{path} not found: build it with `python -m locomanipulationrl_b200.build` "
                      "(there is no fallback path)")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.lm_abi_version() != 1:
        raise LmError("liblm_b200.so ABI version mismatch")
    for which, struct in enumerate((LmTaskParams, LmRobotViews, LmTaskBuffers, LmHostStep, LmPdState, LmRunningStats, LmRolloutStep, LmRolloutBuffers, LmPpoBatch, LmNoiseSpec, LmNoiseState, LmCustomController, LmAdamSchedule, LmActionSink, LmHostPolicy)):
        if lib.lm_struct_size(which) != C.sizeof(struct):
            raise LmError(f"struct layout mismatch for {struct.__name__}: C {lib.lm_struct_size(which)} vs ctypes {C.sizeof(struct)}")
    _lib = lib
    return lib


class _NoRange:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_RANGE = _NoRange()
NVTX = bool(os.environ.get("LM_NVTX"))


def nvtx(name: str):
    """NVTX range around a host-side section (LM_NVTX=1; SURVEY §5 tracing): shows up on the nsys / ncu timeline as
    ``lm/<name>``.  A no-op object otherwise — the step loop pays nothing."""
    if not NVTX:
        return _NO_RANGE
    import torch
    return torch.cuda.nvtx.range("lm/" + name)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().lm_last_error()
        raise LmError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), "tensors passed to the C ABI must be contiguous"
    return t.data_ptr()


def task_params(cfg: TaskConfig, num_envs: int, clamp_actions: bool = True, dof_index=None) -> LmTaskParams:
    """TaskConfig -> LmTaskParams.  Range terms are formed in double like the reference's Python floats
    (utils/math.py:187-189: ``(max_roll-min_roll) * rand``), then rounded to fp32 by ctypes."""
    import math
    p = LmTaskParams()
    p.kind, p.num_envs, p.num_states = cfg.kind, num_envs, cfg.num_states
    p.control_mode, p.zero_actions = cfg.control_mode, int(cfg.zero_actions)
    p.joint1_mirrored = int(cfg.joint1_mirrored)
    p.max_consecutive_successes, p.max_episode_length = cfg.max_consecutive_successes, cfg.max_episode_length
    p.max_reset_counts, p.clamp_actions = cfg.max_reset_counts, int(clamp_actions)
    for j, v in enumerate(dof_index if dof_index is not None else range(NUM_ACT)):
        p.dof_index[j] = int(v)
    p.min_roll, p.roll_range = cfg.min_roll, cfg.max_roll - cfg.min_roll
    p.min_pitch, p.pitch_range = cfg.min_pitch, cfg.max_pitch - cfg.min_pitch
    p.min_yaw, p.yaw_range = cfg.min_yaw, cfg.max_yaw - cfg.min_yaw
    for name in ("position_scale", "quaternion_obs_scale", "linear_vel_scale", "angular_vel_scale", "tip_position_scale",
                 "joint_position_scale", "joint_velocity_scale", "quaternion_scale", "rot_eps", "translation_scale",
                 "joint_acc_scale", "action_rate_scale", "consecutive_success_reward", "fall_penalty", "joint_limit_penalty",
                 "baseline_knee_height", "baseline_height", "baseline_corner_height", "success_thresh",
                 "min_joint_23_diff", "max_joint_23_diff", "reset_min_joint_23_diff", "reset_max_joint_23_diff",
                 "min_joint_1_pos", "max_joint_1_pos", "reset_min_joint_1_pos", "reset_max_joint_1_pos",
                 "clip_actions", "clip_observations"):
        setattr(p, name, getattr(cfg, name))
    p.time_interval = cfg.time_interval
    p.action_limit = {CONTROL_VELOCITY: cfg.velocity_limit, CONTROL_POSITION: math.pi,
                      CONTROL_EFFORT: cfg.torque_limit}.get(cfg.control_mode, cfg.velocity_limit)
    for k in range(4):
        for i in range(3):
            p.corners[k][i] = cfg.corners[k][i]
            p.default_base_tip[k][i] = DEFAULT_BASE_TIP_POSITIONS[k][i]
    for j in range(NUM_DOF):
        p.loco_init_joint_pos[j] = cfg.loco_init_joint_pos[j]
        p.mani_init_joint_pos[j] = cfg.mani_init_joint_pos[j]
    for i in range(3):
        p.loco_default_position[i] = cfg.loco_default_position[i]
        p.mani_default_position[i] = cfg.mani_default_position[i]
        p.obj_default_position[i] = cfg.obj_default_position[i]
        p.indicator_loco_position[i] = cfg.indicator_loco_position[i]
        p.indicator_mani_position[i] = cfg.indicator_mani_position[i]
    for i in range(4):
        p.loco_default_quaternion[i] = cfg.loco_default_quaternion[i]
        p.mani_default_quaternion[i] = cfg.mani_default_quaternion[i]
    return p


class HostBuffer:
    """Page-locked host memory from the library (``lm_host_alloc``), viewed as a CPU float32 tensor.  ``write_combined``: for
    buffers the host only writes (physics state on its way to the device)."""

    def __init__(self, numel: int, write_combined: bool = False):
        import torch
        self._lib = load()
        self.ptr = self._lib.lm_host_alloc(int(numel) * 4, int(bool(write_combined)))
        if not self.ptr:
            raise LmError("lm_host_alloc failed: " + self._lib.lm_last_error().decode())
        self._raw = (C.c_float * int(numel)).from_address(self.ptr)
        self.tensor = torch.frombuffer(self._raw, dtype=torch.float32)

    def data_ptr(self) -> int:
        return self.ptr

    def numel(self) -> int:
        return self.tensor.numel()

    def __del__(self):
        try:
            if self.ptr:
                self._lib.lm_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass
