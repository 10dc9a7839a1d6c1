"""Task constants for every hot-path task variant.

The reference keeps these as class attributes (not yaml), one copy per task file.  Here they are
one dataclass; each named constructor cites the reference class it restates.  The dataclass maps
1:1 onto the C struct ``LmTaskParams`` (include/lm_b200.h) consumed by the CUDA kernels.

Reference shorthand (all under /root/reference/RobotLearning/omniisaacgymenvs/):
  QPC  tasks/quadruped_pose_control_tasks/quadruped_pose_control.py
  QPCV tasks/quadruped_pose_control_tasks/quadruped_pose_control_vertical.py
  QMP  tasks/quadruped_manipulate_plate/quadruped_manipulate_plate.py
  QMPV tasks/quadruped_manipulate_plate/quadruped_manipulate_plate_vertical.py
  JL   tasks/joint_train_locomanipulation/joint_locomanipulation.py
  JLV  tasks/joint_train_locomanipulation/joint_locomanipulation_vertical.py
  *_PC the ``_position_control.py`` sibling of QPC / QMP / JL
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Dict, List, Tuple

KIND_LOCO = 0
KIND_MANI = 1
KIND_COTRAIN = 2

CONTROL_VELOCITY = 0   # robot.py:452-454  (default of all six robot configs, quadruped_robot.py:12,42,77,...)
CONTROL_POSITION = 1   # robot.py:449-451
CONTROL_EFFORT = 2     # robot.py:455-459
CONTROL_PD = 3         # *_PC explicit PD torque loop (quadruped_pose_control_position_control.py:275-313)

# robot/quadruped_robot.py:45-52 (QuadrupedRobotOVOmni / OVFixedBaseOmni)
_HORIZONTAL_ROBOT_INIT = [-1.2, 1.2, 1.2, -1.2,
                          -1.22, -1.92, 1.92, 1.22, 1.92, 1.22, -1.22, -1.92,
                          0.953, -0.953, 0.953, -0.953, 0.953, -0.953, 0.953, -0.953]
# QPC:94-102 / QMP:94-102 override
_HORIZONTAL_TASK_INIT = [-1.57, 1.57, 1.57, -1.57,
                         -1.04, -2.09, 2.09, 1.04, 2.09, 1.04, -1.04, -2.09,
                         1.37, -1.37, 1.37, -1.37, 1.37, -1.37, 1.37, -1.37]
# robot/quadruped_robot.py:80-87 (QuadrupedRobotVerticalOVOmni / VerticalOVFixedOmni)
_VERTICAL_INIT = [0.0, 0.0, 0.0, 0.0,
                  0.35, -0.35, 0.35, -0.35, 0.35, -0.35, 0.35, -0.35,
                  0.95, -0.95, 0.95, -0.95, 0.95, -0.95, 0.95, -0.95]

# QPC:132-137
_HORIZONTAL_CORNERS = [[0.075, 0.1835, -0.04], [-0.075, 0.1835, -0.04],
                       [0.075, -0.1835, -0.04], [-0.075, -0.1835, -0.04]]
# JLV:147-152
_VERTICAL_CORNERS = [[0.0, -0.115, -0.1853], [0.0, 0.115, -0.1853],
                     [-0.115, 0.0, -0.1853], [0.115, 0.0, -0.1853]]

# QPC:127-130 (same in every task file)
DEFAULT_BASE_TIP_POSITIONS = [[-0.0937, 0.1223, -0.1774], [0.0937, 0.1408, -0.1773],
                              [-0.0937, -0.1408, -0.1773], [0.0937, -0.1223, -0.1774]]

# *_PC: quadruped_pose_control_position_control.py:32-52,186-190
_PC_MIN_SE = [-2.35, -0.78, -0.78, -2.35, -2.09, 0.52, 1.05, 0.52, 1.05, 0.52, -2.09, 0.52]
_PC_MAX_SE = [0.78, 2.35, 2.35, 0.78, -1.05, 2.09, 2.09, 2.09, 2.09, 2.09, -1.05, 2.09]
_PC_INIT_SE = [-1.57, 1.57, 1.57, -1.57, -1.57, 1.05, 1.57, 1.05, 1.57, 1.05, -1.57, 1.05]


@dataclass
class TaskConfig:
    name: str
    kind: int
    vertical: bool
    # --- goal sampling (rand_quaternions ranges) QPC:38-43
    min_roll: float = -0.4
    max_roll: float = 0.4
    min_pitch: float = -0.4
    max_pitch: float = 0.4
    min_yaw: float = -1.57
    max_yaw: float = 1.57
    # --- observation scales QPC:27-34
    position_scale: float = 5.0
    quaternion_obs_scale: float = 1.0
    linear_vel_scale: float = 2.0
    angular_vel_scale: float = 0.25
    tip_position_scale: float = 1.0
    joint_position_scale: float = 0.3
    joint_velocity_scale: float = 0.3
    # --- reward QPC:54-61,457
    quaternion_scale: float = 0.5
    rot_eps: float = 0.1
    translation_scale: float = -2.5
    joint_acc_scale: float = -0.0005
    action_rate_scale: float = -0.02
    consecutive_success_reward: float = 600.0
    fall_penalty: float = 0.0
    joint_limit_penalty: float = -5.0
    # --- thresholds QPC:64-87
    baseline_knee_height: float = 0.04
    baseline_height: float = 0.05
    baseline_corner_height: float = 0.01
    success_thresh: float = 0.15
    max_consecutive_successes: int = 15
    min_joint_23_diff: float = 0.43
    max_joint_23_diff: float = 2.53
    reset_min_joint_23_diff: float = 0.384
    reset_max_joint_23_diff: float = 2.61
    min_joint_1_pos: float = -2.35
    max_joint_1_pos: float = 0.78
    reset_min_joint_1_pos: float = -2.44
    reset_max_joint_1_pos: float = 0.87
    joint1_mirrored: bool = True   # horizontal: a2,a3 use (-max,-min) (QPC:482-503); vertical: all four share one range (JLV:589-596)
    corners: List[List[float]] = field(default_factory=lambda: [list(c) for c in _HORIZONTAL_CORNERS])
    # --- defaults written by reset_idx
    loco_init_joint_pos: List[float] = field(default_factory=lambda: list(_HORIZONTAL_TASK_INIT))
    loco_default_position: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.14])
    loco_default_quaternion: List[float] = field(default_factory=lambda: [1.0, 0.0, 0.0, 0.0])
    mani_init_joint_pos: List[float] = field(default_factory=lambda: list(_HORIZONTAL_TASK_INIT))
    mani_default_position: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    mani_default_quaternion: List[float] = field(default_factory=lambda: [0.0, 1.0, 0.0, 0.0])
    obj_default_position: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.14])
    indicator_loco_position: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.3])
    indicator_mani_position: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.3])
    # --- buffers / surface
    num_observations: int = 64
    num_states: int = 93           # 93-wide states_buf in single-task horizontal/vertical files (QPC:119,406-423); 64 (alias of obs) in co-train (JLV:130,515)
    num_actions: int = 12
    zero_actions: bool = False     # debug leftover ``actions[:] = 0`` shipped live in JLV:258-259, QPCV:204, QMPV:214, *_PC
    # --- control
    control_mode: int = CONTROL_VELOCITY
    velocity_limit: float = 3.0    # quadruped_robot.py:55
    torque_limit: float = 1.5      # quadruped_robot.py:56
    # --- sim cfg (cfg/task/*.yaml)
    dt: float = 0.0083
    control_frequency_inv: int = 4
    max_episode_length: int = 300
    clip_observations: float = 5.0
    clip_actions: float = 1.0
    env_spacing: float = 1.0
    max_reset_counts: int = 2048   # QPC:151
    # --- *_PC explicit PD (quadruped_pose_control_position_control.py:25-52)
    pd_kp: float = 4.5
    pd_kd: float = 0.2
    pd_max_effort: float = 1.5
    pd_action_scale: float = 0.1
    pd_control_decimal: int = 4
    pd_torque_divisor: float = 4.0   # single-task: /control_decimal (:313); co-train: /(control_decimal+1) (JL_PC:408)
    pd_min_se: List[float] = field(default_factory=lambda: list(_PC_MIN_SE))
    pd_max_se: List[float] = field(default_factory=lambda: list(_PC_MAX_SE))
    pd_init_se: List[float] = field(default_factory=lambda: list(_PC_INIT_SE))

    pd_sum_final_torque: bool = True   # single-task files add the torque computed after the last sub-step to torque_sum (:311); JL_PC does not (:399-408)
    pd_carry_post_state: bool = True   # single-task files re-point joint_positions_* at the post-step state in get_observations (:392-393); JL_PC keeps the last sub-step's
    pd_target_obs_scale: float = 0.3  # literal ``* 0.3`` on the target rows of the observation (…position_control.py:438-439)
    # --- *_custom_controller.py (quadruped_pose_control_custom_controller.py:24-65): the PD loop of *_PC plus an 88-wide observation
    # (actions AND target rows, :441-455), an unflipped quat_diff (:430), |a| as action-rate term (:531) and three more reward terms
    custom_controller: bool = False
    mechanical_power_penalty_scale: float = -0.02          # :48, :534-535
    position_target_error_penalty_scale: float = -0.05     # :49, :538-539
    rot_dist_decreasing_reward_scale: float = 0.0          # :50 (``# 50``), :542-545
    no_rot_dist_decreasing_reward_thresh: float = 0.3      # :51
    clip_reward: bool = False                              # :65, :619-620
    cc_rot_dist_term: bool = True                          # the rot-dist-decrease term and last_rot_dist exist in the locomotion file only
    cc_update_last_targets: bool = True                    # :724; the manipulation file never refreshes last_joint_position_targets after a reset

    @property
    def time_interval(self) -> float:
        """Finite-difference Δt of ``RobotOmni.update_joint_states`` (robot.py:290).  Velocity-target files hand the robot
        ``dt * controlFrequencyInv`` (QPC:156,189); the *_PC files compute ``_time_inv = dt*(control_decimal+1)`` (:199)
        but hand the robot the bare ``dt`` (:232), and refresh the joint state after every PD sub-step."""
        if self.control_mode == CONTROL_PD:
            return self.dt
        return self.dt * self.control_frequency_inv

    def with_(self, **kw) -> "TaskConfig":
        return replace(self, **kw)


def _vertical_limits() -> Dict[str, object]:
    # JLV:96-99 / QPCV:84-87
    return dict(min_joint_1_pos=-2.09, max_joint_1_pos=2.09, reset_min_joint_1_pos=-2.26,
                reset_max_joint_1_pos=2.26, joint1_mirrored=False,
                corners=[list(c) for c in _VERTICAL_CORNERS])


def quadruped_pose_control() -> TaskConfig:
    """QPC:20-158."""
    return TaskConfig(name="QuadrupedPoseControl", kind=KIND_LOCO, vertical=False)


def quadruped_pose_control_vertical() -> TaskConfig:
    """QPCV: robot defaults from QuadrupedRobotVerticalOVOmni (quadruped_robot.py:70-103); ``actions[:] = 0`` live at :204."""
    return TaskConfig(name="QuadrupedPoseControlVertical", kind=KIND_LOCO, vertical=True,
                      loco_init_joint_pos=list(_VERTICAL_INIT), loco_default_position=[0.0, 0.0, 0.35],
                      indicator_loco_position=[0.0, 0.0, 0.4], zero_actions=True, **_vertical_limits())


def quadruped_manipulate_plate() -> TaskConfig:
    """QMP:21-168."""
    return TaskConfig(name="QuadrupedManipulatePlate", kind=KIND_MANI, vertical=False)


def quadruped_manipulate_plate_vertical() -> TaskConfig:
    """QMPV:91-93,141-143; ``actions[:] = 0`` live at :214."""
    return TaskConfig(name="QuadrupedManipulatePlateVertical", kind=KIND_MANI, vertical=True,
                      mani_init_joint_pos=list(_VERTICAL_INIT), obj_default_position=[0.0, 0.0, 0.35],
                      indicator_mani_position=[0.0, 0.0, 0.4], zero_actions=True, **_vertical_limits())


def joint_locomanipulation() -> TaskConfig:
    """JL:21-227.  Ships fixed goal angles (JL:61-66) and the robot-file default poses (overrides commented out, JL:114-137)."""
    return TaskConfig(name="JointLocomanipulation", kind=KIND_COTRAIN, vertical=False,
                      min_roll=0.2, max_roll=0.2, min_pitch=0.2, max_pitch=0.2, min_yaw=0.785, max_yaw=0.785,
                      loco_init_joint_pos=list(_HORIZONTAL_ROBOT_INIT), loco_default_position=[0.0, 0.0, 0.18],
                      mani_init_joint_pos=list(_HORIZONTAL_ROBOT_INIT), mani_default_position=[0.0, 0.0, 0.5],
                      obj_default_position=[0.0, 0.0, 0.68], indicator_loco_position=[0.0, 0.0, 0.3],
                      indicator_mani_position=[0.0, 0.0, 0.8], num_states=64, env_spacing=0.5)


def joint_locomanipulation_vertical() -> TaskConfig:
    """JLV:21-195; ``actions[...] = 0`` live at JLV:258-259."""
    return TaskConfig(name="JointLocomanipulationVertical", kind=KIND_COTRAIN, vertical=True,
                      loco_init_joint_pos=list(_VERTICAL_INIT), loco_default_position=[0.0, 0.0, 0.35],
                      mani_init_joint_pos=list(_VERTICAL_INIT), mani_default_position=[0.0, 0.0, 0.6],
                      obj_default_position=[0.0, 0.0, 0.95], indicator_loco_position=[0.0, 0.0, 0.5],
                      indicator_mani_position=[0.0, 0.0, 1.1], num_states=64, zero_actions=True,
                      **_vertical_limits())


def _position_control(**kw) -> Dict[str, object]:
    """Deltas shared by the three *_position_control.py files: effort mode + explicit PD loop, 64-wide states (alias of obs),
    success threshold 0.1, the horizontal task pose."""
    d = dict(vertical=False, control_mode=CONTROL_PD, num_states=64, success_thresh=0.1, control_frequency_inv=1,
             loco_init_joint_pos=list(_HORIZONTAL_TASK_INIT), mani_init_joint_pos=list(_HORIZONTAL_TASK_INIT))
    d.update(kw)
    return d


def quadruped_pose_control_position_control() -> TaskConfig:
    """quadruped_pose_control_position_control.py:20-199; ``actions[:] = 0.0`` live at :261; cfg/task/QuadrupedPoseControlPositionControl.yaml."""
    return TaskConfig(name="QuadrupedPoseControlPositionControl", kind=KIND_LOCO,
                      **_position_control(min_roll=-0.5, max_roll=0.5, min_pitch=-0.5, max_pitch=0.5, min_yaw=-3.14, max_yaw=3.14,
                                          zero_actions=True, dt=0.005, max_episode_length=500,
                                          pd_control_decimal=4, pd_torque_divisor=4.0))


def quadruped_manipulate_plate_position_control() -> TaskConfig:
    """quadruped_manipulate_plate_position_control.py; ``actions[:] = 0.0`` live at :274; robot at z 0.3, plate at 0.44."""
    return TaskConfig(name="QuadrupedManipulatePlatePositionControl", kind=KIND_MANI,
                      **_position_control(min_roll=-0.5, max_roll=0.5, min_pitch=-0.5, max_pitch=0.5, min_yaw=-3.14, max_yaw=3.14,
                                          zero_actions=True, dt=0.005, max_episode_length=450,
                                          mani_default_position=[0.0, 0.0, 0.3], obj_default_position=[0.0, 0.0, 0.44],
                                          indicator_mani_position=[0.0, 0.0, 0.5], pd_control_decimal=4, pd_torque_divisor=4.0))


def joint_locomanipulation_position_control() -> TaskConfig:
    """joint_locomanipulation_position_control.py:20-408: control_decimal 7, torque mean over decimal + 1 (:408), no action zeroing."""
    return TaskConfig(name="JointLocomanipulationPositionControl", kind=KIND_COTRAIN,
                      **_position_control(joint_acc_scale=-0.001, action_rate_scale=-0.03, dt=0.00417, max_episode_length=300,
                                          loco_default_position=[0.0, 0.0, 0.14], mani_default_position=[0.0, 0.0, 0.5],
                                          obj_default_position=[0.0, 0.0, 0.64], indicator_loco_position=[0.0, 0.0, 0.3],
                                          indicator_mani_position=[0.0, 0.0, 0.8], pd_control_decimal=7, pd_torque_divisor=8.0,
                                          pd_sum_final_torque=False, pd_carry_post_state=False))


_CC_INIT_SE = [-1.2, 1.2, 1.2, -1.2, -1.57, 0.7, 1.57, 0.7, 1.57, 0.7, -1.57, 0.7]   # quadruped_pose_control_custom_controller.py:180-184


def quadruped_pose_control_custom_controller() -> TaskConfig:
    """quadruped_pose_control_custom_controller.py:20-190; cfg/task/QuadrupedPoseControlCustomController.yaml.  Robot-file defaults
    (no pose override in __init__, :124-126), fixed goal yaw 1.57 (:38-43), 93-wide states (:161), actions are NOT zeroed (:258)."""
    return TaskConfig(name="QuadrupedPoseControlCustomController", kind=KIND_LOCO, vertical=False, control_mode=CONTROL_PD,
                      custom_controller=True, num_observations=88, num_states=93,
                      min_roll=-0.0, max_roll=0.0, min_pitch=-0.0, max_pitch=0.0, min_yaw=1.57, max_yaw=1.57,
                      joint_acc_scale=-0.00015, action_rate_scale=-0.01, success_thresh=0.15, max_consecutive_successes=20,
                      loco_init_joint_pos=list(_HORIZONTAL_ROBOT_INIT), loco_default_position=[0.0, 0.0, 0.18],
                      dt=0.005, control_frequency_inv=1, max_episode_length=500,
                      pd_control_decimal=4, pd_torque_divisor=4.0, pd_init_se=list(_CC_INIT_SE))


def quadruped_manipulate_plate_custom_controller() -> TaskConfig:
    """quadruped_manipulate_plate_custom_controller.py; cfg/task/QuadrupedManipulatePlateCustomController.yaml.  Robot at the
    origin with the robot-file default pose, plate at z 0.18, no rot-dist-decrease term, last targets frozen after a reset."""
    return TaskConfig(name="QuadrupedManipulatePlateCustomController", kind=KIND_MANI, vertical=False, control_mode=CONTROL_PD,
                      custom_controller=True, cc_rot_dist_term=False, cc_update_last_targets=False, num_observations=88, num_states=93,
                      joint_acc_scale=-0.00015, action_rate_scale=-0.01, success_thresh=0.15, max_consecutive_successes=20,
                      mani_init_joint_pos=list(_HORIZONTAL_ROBOT_INIT), mani_default_position=[0.0, 0.0, 0.0],
                      obj_default_position=[0.0, 0.0, 0.18], indicator_mani_position=[0.0, 0.0, 0.3],
                      dt=0.005, control_frequency_inv=1, max_episode_length=500,
                      pd_control_decimal=4, pd_torque_divisor=4.0, pd_init_se=list(_CC_INIT_SE))


TASK_CONFIGS = {
    "QuadrupedPoseControl": quadruped_pose_control,
    "QuadrupedPoseControlVertical": quadruped_pose_control_vertical,
    "QuadrupedManipulatePlate": quadruped_manipulate_plate,
    "QuadrupedManipulatePlateVertical": quadruped_manipulate_plate_vertical,
    "JointLocomanipulation": joint_locomanipulation,
    "JointLocomanipulationVertical": joint_locomanipulation_vertical,
    "QuadrupedPoseControlPositionControl": quadruped_pose_control_position_control,
    "QuadrupedManipulatePlatePositionControl": quadruped_manipulate_plate_position_control,
    "JointLocomanipulationPositionControl": joint_locomanipulation_position_control,
    "QuadrupedPoseControlCustomController": quadruped_pose_control_custom_controller,
    "QuadrupedManipulatePlateCustomController": quadruped_manipulate_plate_custom_controller,
}


def get_task_config(name: str) -> TaskConfig:
    """Same role as the reference's ``task_map`` lookup (utils/task_util.py:60-90)."""
    if name not in TASK_CONFIGS:
        raise KeyError(f"unknown task {name!r}; known: {sorted(TASK_CONFIGS)}")
    return TASK_CONFIGS[name]()
