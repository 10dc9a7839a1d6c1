/*
 * lm_b200.h — C ABI of the B200-native LocoManipulationRL task / policy hot path.
 *
 * Drop-in boundary for the reference's per-environment task math.  Every entry point cites the
 * reference interface it replaces (paths under /root/reference/RobotLearning/omniisaacgymenvs/).
 * All pointers are DEVICE pointers unless a name ends in _host; all tensors are contiguous,
 * row-major, 16-byte aligned, in the reference's own dtypes (fp32 data, int64 flags).  `stream`
 * is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).  Calls never
 * allocate, never synchronise the host (except the *_host entry points), and never throw:
 * they return 0 on success or a negative LM_ERR_* / positive cudaError_t code;
 * lm_last_error() returns a thread-local message.
 */
#ifndef LM_B200_H
#define LM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LM_ABI_VERSION 1

#define LM_NUM_DOF 20        /* articulation DOFs: 12 actuated + 8 passive (robot/base/robot.py:437) */
#define LM_NUM_ACT 12        /* actions / actuated DOFs (quadruped_pose_control.py:117)              */
#define LM_NUM_OBS 64        /* observation row (quadruped_pose_control.py:118,358-371)              */
#define LM_NUM_STATES_FULL 93 /* single-task states row (quadruped_pose_control.py:119,406-423)      */

enum { LM_KIND_LOCO = 0, LM_KIND_MANI = 1, LM_KIND_COTRAIN = 2 };
enum { LM_CONTROL_VELOCITY = 0, LM_CONTROL_POSITION = 1, LM_CONTROL_EFFORT = 2, LM_CONTROL_PD = 3 };
enum { LM_GAE_RLGAMES = 0, LM_GAE_SKRL = 1 };

enum {
  LM_OK = 0,
  LM_ERR_BAD_ARG = -1,     /* null pointer / inconsistent sizes */
  LM_ERR_ALIGNMENT = -2,   /* a tensor is not 16-byte aligned   */
  LM_ERR_NO_DEVICE = -3,   /* no sm_100 device / wrong arch     */
  LM_ERR_UNSUPPORTED = -4
};

/* Task constants: mirror of the reference's class attributes
 * (quadruped_pose_control.py:27-87, quadruped_manipulate_plate.py:28-88,
 *  joint_locomanipulation_vertical.py:39-99) plus the yaml values the tasks read
 *  (cfg/task/*.yaml: dt, controlFrequencyInv, max_episode_length, clipObservations, clipActions). */
typedef struct LmTaskParams {
  int32_t kind;                 /* LM_KIND_*                                                         */
  int32_t num_envs;             /* N; co-train: first N/2 locomotion, last N/2 manipulation (JL:22-35) */
  int32_t num_states;           /* 93 (single task) or 64 (co-train: states alias obs, JLV:515)      */
  int32_t control_mode;         /* LM_CONTROL_*                                                      */
  int32_t zero_actions;         /* debug leftover `actions[:] = 0` shipped live (JLV:258-259)        */
  int32_t joint1_mirrored;      /* 1: a2,a3 use (-max,-min) (QPC:482-503); 0: one range (JLV:589-596) */
  int32_t max_consecutive_successes;
  int32_t max_episode_length;
  int32_t max_reset_counts;     /* 2048 (QPC:151)                                                    */
  int32_t clamp_actions;        /* 1: fuse VecEnvRLGames.step's clamp (envs/vec_env_rlgames.py:60)   */
  int32_t dof_index[LM_NUM_ACT];/* articulation column of each actuated DOF (robot.py:174-190)       */
  /* goal sampling (utils/math.py:176-193): angle = min + range * u, range = (max - min) in double    */
  float min_roll, roll_range, min_pitch, pitch_range, min_yaw, yaw_range;
  /* observation scales (QPC:27-34) */
  float position_scale, quaternion_obs_scale, linear_vel_scale, angular_vel_scale, tip_position_scale,
      joint_position_scale, joint_velocity_scale;
  /* reward (QPC:54-61,432-457) */
  float quaternion_scale, rot_eps, translation_scale, joint_acc_scale, action_rate_scale,
      consecutive_success_reward, fall_penalty, joint_limit_penalty;
  /* thresholds (QPC:64-87) */
  float baseline_knee_height, baseline_height, baseline_corner_height, success_thresh;
  float min_joint_23_diff, max_joint_23_diff, reset_min_joint_23_diff, reset_max_joint_23_diff;
  float min_joint_1_pos, max_joint_1_pos, reset_min_joint_1_pos, reset_max_joint_1_pos;
  float time_interval;          /* dt * controlFrequencyInv (QPC:156)                                 */
  float clip_actions, clip_observations;
  float action_limit;           /* velocity ±3 rad/s, position ±pi, effort ±1.5 N·m (robot.py:195-231) */
  float corners[4][3];          /* body corners in the robot frame (QPC:132-137, JLV:147-152)         */
  float default_base_tip[4][3]; /* QPC:127-130                                                        */
  float loco_init_joint_pos[LM_NUM_DOF], loco_default_position[3], loco_default_quaternion[4];
  float mani_init_joint_pos[LM_NUM_DOF], mani_default_position[3], mani_default_quaternion[4];
  float obj_default_position[3];
  float indicator_loco_position[3], indicator_mani_position[3];
} LmTaskParams;

/* One articulation + its auxiliary rigid views, as PhysX exposes them through
 * ArticulationView / RigidPrimView (robot/base/robot.py:288-321; objects/base/rigid_object.py:52-58).
 * Read by lm_task_post_step; written for reset envs by lm_task_pre_step
 * (the set_* calls of reset_idx, QPC:244-284 / QMP:252-294).  n = N (single task) or N/2. */
typedef struct LmRobotViews {
  float* dof_pos;        /* [n,20] */
  float* dof_vel;        /* [n,20] */
  float* root_pos;       /* [n,3] world (locomotion robot only; NULL for the fixed-base robot) */
  float* root_quat;      /* [n,4] wxyz  */
  float* root_lin_vel;   /* [n,3] */
  float* root_ang_vel;   /* [n,3] */
  float* tip_pos;        /* [n*4,3] world */
  float* knee_pos;       /* [n*8,3] world */
  float* indicator_pos;  /* [n,3] world  (pose indicator, objects/pose_indicator.py) */
  float* indicator_quat; /* [n,4] */
  float* plate_pos;      /* [n,3] world  (manipulation only; objects/plate.py) */
  float* plate_quat;     /* [n,4] */
  float* plate_lin_vel;  /* [n,3] */
  float* plate_ang_vel;  /* [n,3] */
  float* targets;        /* [n,12] joint velocity/position targets, or [n,20] efforts (robot.py:444-461) */
} LmRobotViews;

/* Task-owned RL buffers (tasks/base/rl_task.py:104-113; quadruped_pose_control.py:123-154). */
typedef struct LmTaskBuffers {
  float* obs_buf;                  /* [N,64]                                     */
  float* states_buf;               /* [N,93] or NULL when states alias obs       */
  float* rew_buf;                  /* [N]                                        */
  int64_t* reset_buf;              /* [N] (initialised to ones, rl_task.py:111)  */
  int64_t* progress_buf;           /* [N]                                        */
  int64_t* successes;              /* [N]                                        */
  int64_t* consecutive_successes;  /* [N]                                        */
  int64_t* goal_reset_buf;         /* [N]                                        */
  float* actions_ring;             /* [2,N,12]: slot (step&1) = current_actions, other = last_actions */
  float* last_joint_vel;           /* [N,12] (robot.py:263,291)                  */
  float* last_base_tip;            /* [N,4,3] (QPC:126,426)                      */
  float* goal_quaternions;         /* [N,4]                                      */
  int64_t* reset_env_ids;          /* [N] out: ascending ids reset by the last pre_step (QPC:206)      */
  int32_t* num_reset;              /* [2] out: {count, count in locomotion half}                        */
  float* extras;                   /* [LM_NUM_EXTRAS] out, see LM_EXTRA_*                               */
  int64_t* counters;               /* [6] num_successes,num_resets,{_loco},{_mani} (QPC:152-153, JLV:173-184) */
} LmTaskBuffers;

enum {
  LM_EXTRA_ORIENTATION_REW = 0, LM_EXTRA_TRANSLATION_PENALTY, LM_EXTRA_JOINT_ACC_PENALTY,
  LM_EXTRA_ACTION_RATE_PENALTY, LM_EXTRA_CONSECUTIVE_SUCCESSES_REW, LM_EXTRA_JOINT_LIMIT_PENALTY,
  LM_EXTRA_FALL_PENALTY, LM_EXTRA_SUCCESS_RATE, LM_EXTRA_SUCCESS_RATE_LOCO, LM_EXTRA_SUCCESS_RATE_MANI,
  LM_EXTRA_NAN_COUNT, LM_EXTRA_BELOW_M50_COUNT, /* device-side form of the guards at QPC:550-558 */
  LM_NUM_EXTRAS = 16
};

/* State of the explicit position-PD variants (*_position_control.py: quadruped_pose_control_position_control.py:179-191,
 * joint_locomanipulation_position_control.py:161-183), env-indexed like the RL buffers ([N,12]; co-train: locomotion rows
 * first).  Registered once with lm_task_set_pd_state; then lm_task_pre_step resets these rows for reset envs
 * (…position_control.py:377-386) and lm_task_post_step writes 0.3 * targets / 0.3 * last_targets into obs[40:64] instead
 * of the current / last actions (…:438-452).  As in the reference, last_targets is only ever written on reset. */
typedef struct LmPdState {
  float* joint_pos;       /* [N,12] joint positions carried between PD sub-steps                 */
  float* joint_vel;       /* [N,12]                                                              */
  float* targets;         /* [N,12] current_joint_position_targets                               */
  float* last_targets;    /* [N,12] last_joint_position_targets                                  */
  float* targets_se;      /* [N,12] swing / extension form                                       */
  float init_se[LM_NUM_ACT]; /* reset value of targets_se                                        */
  int32_t carry_post_state; /* 1: post_step stores the observed joint state into joint_pos / joint_vel
                             *    (single-task files, …position_control.py:392-393); 0: JL_PC            */
  float target_obs_scale; /* the literal 0.3 of …position_control.py:438-439                     */
} LmPdState;

typedef struct LmTask LmTask; /* opaque */

int lm_abi_version(void);
const char* lm_last_error(void);
/* number of SMs / name of the device the library will launch on (-1 / "" if none) */
int lm_device_sm_count(int device);
/* sizeof() of the ABI structs as compiled (0 LmTaskParams, 1 LmRobotViews, 2 LmTaskBuffers, 3 LmHostStep, 4 LmPdState):
 * bindings assert their mirror structs agree. */
int lm_struct_size(int which);

/* Replaces the constructor + cleanup() of RLTask/JointRLTask (rl_task.py:55-113): validates
 * params, sizes the launch grids, allocates the few KB of device scratch (tile reset counts,
 * block partials).  The only allocating call. */
LmTask* lm_task_create(const LmTaskParams* params, int device);
void lm_task_destroy(LmTask* task);
/* Current step parity (which actions_ring slot holds current_actions). */
int lm_task_step_parity(const LmTask* task);
/* When on, post_step stores clamp(obs, ±clip_observations) — the adapter's _process_data clamp
 * (envs/vec_env_rlgames.py:42) fused into the producing kernel.  Off by default (obs_buf then holds
 * the reference task's unclamped values). */
int lm_task_set_clamp_observations(LmTask* task, int on);
/* post_step leaves behind the exclusive scan of the reset mask it just produced, so the next pre_step needs no
 * extra pass over reset_buf.  Call this after writing reset_buf from outside the library (e.g. RLTask.reset(),
 * rl_task.py:222-225): the next pre_step then rescans reset_buf first. */
int lm_task_invalidate_reset_scan(LmTask* task);

/* Diagnostic: when stamps_dev != NULL (device buffer of num_tiles x 8 int64, num_tiles = ceil(envs per half / 128) per
 * half), every CTA of the following lm_task_post_step launches records %globaltimer at its stages.  NULL = off (default). */
int lm_task_set_timeline(LmTask* task, long long* stamps_dev);

/* Registers (or, with NULL, clears) the PD state of a LM_CONTROL_PD task. */
int lm_task_set_pd_state(LmTask* task, const LmPdState* pd);

/* pre_physics_step (QPC:200-228 / QMP:213-236 / JLV:248-270) including reset_idx
 * (QPC:230-299 / JLV:272-377), rand_quaternions (utils/math.py:176-193) and take_action
 * (robot.py:444-461).  `actions` [N,12]; `uniforms` [N,3] = the torch.rand draw, row k is consumed
 * by the k-th reset env in ascending id order (its reset rank / goal index).  `loco` / `mani`
 * may be NULL when the task kind has no such half. */
int lm_task_pre_step(LmTask* task, const LmRobotViews* loco, const LmRobotViews* mani,
                     const float* env_pos, const LmTaskBuffers* buf, const float* actions,
                     const float* uniforms, void* stream);

/* reset_idx(env_ids) called directly (QPC:230-299): env_ids [count] int64 (device), entry k consumes
 * uniform row k.  In co-train the caller orders ids locomotion-first, as nonzero() does. */
int lm_task_reset_idx(LmTask* task, const LmRobotViews* loco, const LmRobotViews* mani,
                      const float* env_pos, const LmTaskBuffers* buf, const int64_t* env_ids,
                      int32_t count, const float* uniforms, void* stream);

/* post_physics_step (rl_task.py:240-260): progress += 1, update_all_states (robot.py:276-321),
 * get_observations (QPC:301-426 / QMP:311-435 / JLV:379-521), calculate_metrics
 * (QPC:428-560), is_done (QPC:562-633) and the extras means / success-rate bookkeeping. */
int lm_task_post_step(LmTask* task, const LmRobotViews* loco, const LmRobotViews* mani,
                      const float* env_pos, const LmTaskBuffers* buf, void* stream);

/* Action sink: the action half of pre_physics_step (adapter clamp, current_actions stash, take_action) handed to the kernel that
 * produces the actions.  lm_task_action_sink fills the struct for the NEXT pre_physics_step (call it after lm_task_post_step);
 * lm_mlp_forward_sink (evaluation: action = policy mean) and lm_rollout_write (LmRolloutStep.sink: action = the sample) then
 * deliver the actions into the task's buffers themselves, and the next lm_task_pre_step is called with actions == NULL: it runs
 * only the reset half (reset scan, reset_idx scatter, goal resample).  Velocity / position control modes (the effort mode's
 * 20-wide scatter and the explicit-PD files keep the plain path). */
typedef struct LmActionSink {
  float* ring_next;      /* [N,12] the slot of actions_ring the next pre_physics_step treats as current_actions */
  float* targets_loco;   /* [H,12] or NULL */
  float* targets_mani;   /* [H,12] or NULL */
  int32_t loco_envs;     /* envs [0, loco_envs) drive the locomotion robot's targets, the rest the manipulation robot's */
  int32_t clamp_actions; int32_t zero_actions; int32_t pad_;
  float clip_actions; float action_limit;
} LmActionSink;
/* Deferred finish pass (opt-in pipelining, see lm_task.cu): with defer on, lm_task_post_step only publishes per-tile partial sums and
 * reset counts (the post kernel sheds its serial last-CTA tail); extras / counters / num_reset are then produced by ONE extra CTA
 * inside the next lm_task_pre_step's launch, beside the reset bodies, or by lm_task_post_finish when they are wanted earlier (e.g. at
 * a rollout boundary).  Until one of the two has run they describe the step before.  The pass acts at most once per post step, so
 * calling lm_task_post_finish in front of a captured step graph that carries the extra CTA is safe. */
int lm_task_set_defer_finish(LmTask* task, int on);
/* Early reset (opt-in pipelining, needs an action sink): lm_task_pre_step(actions == NULL) then runs BESIDE the kernel launched
 * before it on the stream (programmatic dependent launch; it completes after that kernel).  That kernel must not touch what the
 * reset half reads / writes and must trigger its dependents only after its own dependency wait (lm_mlp_forward_sink qualifies;
 * lm_rollout_write does NOT: it reads reset_buf as the step's dones, which the reset half clears). */
int lm_task_set_early_reset(LmTask* task, int on);
int lm_task_post_finish(LmTask* task, const LmTaskBuffers* buf, void* stream);
int lm_task_set_step_parity(LmTask* task, int parity);   /* after replaying a captured step: the ring slot (lm_task_step_parity at capture) it used */
int lm_task_action_sink(LmTask* task, const LmRobotViews* loco, const LmRobotViews* mani, const LmTaskBuffers* buf, LmActionSink* out);

/* _process_data of the vec-env adapter (envs/vec_env_rlgames.py:41-46): out = clamp(in, ±clip). */
int lm_clamp_copy(const float* in, float* out, int64_t count, float clip, void* stream);

/* One sub-step of the *_PC explicit PD loop
 * (quadruped_pose_control_position_control.py:295-311): torque = clamp(kp (target - q) - kd qd, ±max),
 * torque_sum += torque.  q/qd are [n,20] articulation rows gathered through dof_index. */
int lm_pd_substep(const float* dof_pos, const float* dof_vel, const float* targets, float* torque,
                  float* torque_sum, const int32_t* dof_index_host, int32_t n, float kp, float kd,
                  float max_effort, void* stream);
/* The same torque law on the CARRIED joint state [n,12] (the first torque of a step uses the state read after the
 * previous step's last sub-step, …position_control.py:295-297). */
int lm_pd_torque(const float* joint_pos, const float* joint_vel, const float* targets, float* torque,
                 float* torque_sum, int32_t n, float kp, float kd, float max_effort, void* stream);
/* update_joint_states (robot.py:283-291) after a physics sub-step, fused with the next torque: gathers the actuated
 * DOFs of the articulation rows into joint_pos / joint_vel, sets last_joint_vel = joint_vel, then applies the torque
 * law.  accumulate = 0 leaves torque_sum alone (joint_locomanipulation_position_control.py:399-408 does not add the
 * torque computed after the last sub-step).  With mean_torque != NULL (the last call of a step) it also stores
 * torque_sum / torque_divisor, the `torque_sum / control_decimal` of …position_control.py:313. */
int lm_pd_update_and_torque(const float* dof_pos, const float* dof_vel, const int32_t* dof_index_host,
                            float* joint_pos, float* joint_vel, float* last_joint_vel, const float* targets,
                            float* torque, float* torque_sum, int32_t accumulate, float* mean_torque,
                            float torque_divisor, int32_t n, float kp, float kd, float max_effort, void* stream);
/* Swing/extension target integration of the *_PC variants (…position_control.py:275-287). */
int lm_pd_integrate_targets(const float* actions, float* targets_se, float* targets,
                            const float* se_lower_host, const float* se_upper_host, int32_t n,
                            float action_scale, int32_t zero_actions, void* stream);

/* MLP actor-critic forward: Linear(64,256)-ELU-Linear(256,128)-ELU-Linear(128,64)-ELU, heads
 * mean Linear(64,12) and value Linear(64,1) (scripts/skrl_ppo_locomanipulation.py:33-42,57-61;
 * cfg/train/JointLocomanipulationPPO.yaml:10-32).  bf16 operands, fp32 accumulate, tcgen05.
 * lm_mlp_pack converts nn.Linear [out,in] fp32 weights (device) into the kernel's resident layout. */
typedef struct LmMlp LmMlp;
LmMlp* lm_mlp_create(int device);
void lm_mlp_destroy(LmMlp* mlp);
int lm_mlp_pack(LmMlp* mlp, const float* w1, const float* b1, const float* w2, const float* b2,
                const float* w3, const float* b3, const float* w_mu, const float* b_mu,
                const float* w_value, const float* b_value, void* stream);
/* obs [N,64] fp32.  Optional input standardisation fused in the prologue: when norm_mean != NULL,
 * x = clamp((obs - norm_mean) * norm_inv_std, ±norm_clip) with norm_inv_std supplied by the caller in
 * its library's form (rl_games RunningMeanStd: 1/sqrt(var + 1e-5), clamp 5; skrl RunningStandardScaler:
 * 1/(sqrt(var) + 1e-8), clamp 5).  Outputs mu [N,12], value [N] fp32. */
int lm_mlp_forward(LmMlp* mlp, const float* obs, const float* norm_mean, const float* norm_inv_std,
                   float norm_clip, float* mu, float* value, int32_t n, void* stream);
/* lm_mlp_forward whose head epilogue also delivers action = mean into the task (LmActionSink above). */
int lm_mlp_forward_sink(LmMlp* mlp, const float* obs, const float* norm_mean, const float* norm_inv_std, float norm_clip, float* mu,
                        float* value, int32_t n, const LmActionSink* sink, void* stream);

/* Diagnostic: when stamps_dev != NULL (device buffer of >= 256 int64), CTA 0 of every following lm_mlp_forward records
 * clock64() at its pipeline stages (see csrc/lm_mlp.cu `stamp`).  NULL switches it off (the default). */
int lm_mlp_set_timeline(LmMlp* mlp, long long* stamps_dev);

/* GNN limb-graph policy forward (scripts/graph_model_orebot_ov.py:11-241; wired at
 * scripts/skrl_ppo_locomanipulation.py:44-49 with hidden 32).  fp32.  Weights are the module's
 * own tensors in nn.Linear layout, concatenated by lm_gnn_pack in a fixed order. */
typedef struct LmGnn LmGnn;
LmGnn* lm_gnn_create(int device);
void lm_gnn_destroy(LmGnn* gnn);
int lm_gnn_pack(LmGnn* gnn, const float* const* weights, int32_t num_tensors, void* stream);
int lm_gnn_forward(LmGnn* gnn, const float* obs, float* mu, float* value, int32_t n, void* stream);
/* The same forward on tcgen05 tensor cores: bf16 operands (weights and node features), fp32 accumulation, fp32 input
 * layers and heads; logits within 1e-2 of lm_gnn_forward (north_star tolerance for bf16 policy logits). */
int lm_gnn_forward_bf16(LmGnn* gnn, const float* obs, float* mu, float* value, int32_t n, void* stream);
/* PPO update of the GNN policy: gradients of a loss w.r.t. all 20 parameter tensors (the order of lm_gnn_pack) from
 * dLoss/dmu [n,12] and dLoss/dvalue [n] (lm_ppo_loss).  One kernel recomputes the fp32 forward per env and walks the limb graph
 * backwards (max-aggregation routes to the first arg-max in-edge); weight gradients accumulate per CTA in shared memory and are
 * summed in CTA order.  workspace: workspace_ctas * lm_gnn_grad_workspace_floats() floats.  Gradients are WRITTEN (not accumulated). */
int32_t lm_gnn_grad_workspace_floats(void);
int lm_gnn_backward(LmGnn* gnn, const float* obs, const float* grad_mu, const float* grad_value, int32_t n, float* workspace,
                    int32_t workspace_ctas, float* const* grads, int32_t num_grads, void* stream);
/* Diagnostic, as lm_mlp_set_timeline: clock64() stamps of CTA 0 of every following lm_gnn_forward_bf16 (>= 512 int64). */
int lm_gnn_set_timeline(LmGnn* gnn, long long* stamps_dev);

/* PPO generalised advantage estimation over a time-major [T,N] rollout.
 * LM_GAE_RLGAMES: rl_games 1.6.0 A2CBase.discount_values (next-step done mask; last step uses
 *   `last_dones`/`last_values`).  LM_GAE_SKRL: skrl PPO compute_gae (current-step done mask;
 *   returns = adv + values; optional normalisation is a separate call).
 * dones is fp32 (0/1) as both libraries store it.  advantages/returns [T,N]. */
int lm_gae(const float* rewards, const float* values, const float* dones, const float* last_values,
           const float* last_dones, float* advantages, float* returns, int32_t T, int32_t N,
           float gamma, float lambda, int32_t variant, void* stream);
/* Same recurrence over an env-major [N,T] buffer: one warp per env, warp-shuffle scan of the affine
 * maps A -> delta_t + c_t A (32 time steps per scan).  Agrees with lm_gae to fp32 round-off. */
int lm_gae_env_major(const float* rewards, const float* values, const float* dones,
                     const float* last_values, const float* last_dones, float* advantages,
                     float* returns, int32_t T, int32_t N, float gamma, float lambda, int32_t variant,
                     void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * *_custom_controller.py task variants (quadruped_pose_control_custom_controller.py and
 * quadruped_manipulate_plate_custom_controller.py; line numbers below are the former's): the explicit PD loop of
 * the *_position_control files plus an 88-wide observation, an unflipped quat_diff, |a| as action-rate term, three more
 * reward terms and the last_rot_dist state.  The task kernels run unchanged (93-wide states variant, action-rate scale 0,
 * observations into a 64-wide scratch); these two calls add the variant's own rows:
 *   lm_cc_pre  after lm_task_pre_step:  last_rot_dist of the envs reset this step (:384-386);
 *   lm_cc_post after lm_task_post_step: obs [N,88] = scratch row with the raw quat_diff + 0.3 * targets, 0.3 * last_targets
 *     (:430-455), raw quat_diff into the states row (:466-479), reward += action-rate + mechanical-power + target-error +
 *     rot-dist-decrease terms (:531-546,608-620), their means into extras[3], [12..14], last_rot_dist <- rot_dist,
 *     last_targets <- targets (:724), carried joint state <- post-step joint state (:392-393). */
enum { LM_CC_UPDATE_LAST_TARGETS = 1, LM_CC_ROT_DIST_TERM = 2 };
enum { LM_EXTRA_MECHANICAL_POWER_PENALTY = 12, LM_EXTRA_TARGET_ERROR_PENALTY = 13, LM_EXTRA_ROT_DIST_DECREASING_REW = 14 };
typedef struct LmCustomController {
  int32_t num_envs; int32_t clip_reward; int32_t clamp_observations;
  int32_t flags;               /* LM_CC_UPDATE_LAST_TARGETS (:724; the manipulation file never does), LM_CC_ROT_DIST_TERM (locomotion file only) */
  float clip_observations; float target_obs_scale;
  float action_rate_scale; float mechanical_power_penalty_scale; float position_target_error_penalty_scale;
  float rot_dist_decreasing_reward_scale; float no_rot_dist_decreasing_reward_thresh; float default_quaternion[4];
  int32_t dof_index[12];
  const float* obs64;          /* [N,64] scratch written by lm_task_post_step          */
  const float* dof_pos;        /* [N,20] post-step articulation state                  */
  const float* dof_vel;        /* [N,20]                                               */
  const float* goal;           /* [N,4]                                                */
  const float* torque;         /* [N,12] mean PD torque of this step (:313)            */
  float* targets;              /* [N,12] current_joint_position_targets                */
  float* last_targets;         /* [N,12]                                               */
  float* joint_pos;            /* [N,12] carried joint state of the PD loop            */
  float* joint_vel;            /* [N,12]                                               */
  float* last_rot_dist;        /* [N]    NULL without LM_CC_ROT_DIST_TERM              */
  float* obs;                  /* [N,88] out                                           */
  float* states;               /* [N,93] patched in place; columns 6..9 hold the frame quaternion k_post computed */
  float* rew;                  /* [N]    in / out                                      */
  float* extras;               /* [LM_NUM_EXTRAS]                                      */
  const int64_t* reset_env_ids; const int32_t* num_reset;   /* lm_cc_pre: what lm_task_pre_step left */
} LmCustomController;
typedef struct LmRollout LmRollout;   /* workspace handle (lm_rollout_create, below) */
int lm_cc_pre(const LmCustomController* cc, void* stream);
int lm_cc_post(LmRollout* workspace, const LmCustomController* cc, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Rollout writer + running mean/std (the caller-side formats either side of the path).  Not in the
 * reference repo: rl_games 1.6.0 ExperienceBuffer / RunningMeanStd / a2c_continuous.neglogp (cfg/train/*PPO.yaml:46-47,
 * 66) and skrl RandomMemory / RunningStandardScaler / GaussianMixin (scripts/skrl_ppo_locomanipulation.py:100-122).
 * Buffers are time-major [T,N,.] so that lm_gae and the PPO update read them coalesced. */
enum { LM_RMS_RLGAMES = 0, LM_RMS_SKRL = 1 };
typedef struct LmRunningStats {
  double* mean;    /* [dim] float64, as both libraries keep them */
  double* var;     /* [dim] */
  double* count;   /* [1]   */
  int32_t dim;     /* 1 or a multiple of 4, <= 64 */
  int32_t variant; /* LM_RMS_RLGAMES: y = (x - mean) / sqrt(var + 1e-5); LM_RMS_SKRL: y = (x - mean) / (sqrt(var) + 1e-8) */
} LmRunningStats;
typedef struct LmRolloutStep {   /* one env step; device pointers */
  const float* obs;       /* [N,64] what the policy saw               */
  const float* mu;        /* [N,12] policy mean                       */
  const float* value;     /* [N]                                      */
  const float* log_std;   /* [12]   state-independent log sigma       */
  const float* noise;     /* [N,12] standard-normal draws of the caller's generator; NULL: see sample_in_kernel */
  const float* rew;       /* [N]    the reward that arrived WITH this observation = the reward of the action in slot t-1 */
  const int64_t* dones;   /* [N]    reset_buf                         */
  /* noise == NULL: sample_in_kernel != 0 draws N(0,1) in the kernel (Philox4x32-10 keyed by noise_seed, stream position noise_step,
   * element index: no draw tensor in HBM, reproducible for a given seed / step); 0: action = mu (evaluation) */
  uint64_t noise_seed; uint64_t noise_step; int32_t sample_in_kernel; int32_t pad_;
  /* rl_games value_bootstrap (cfg/train/*PPO.yaml:48): rewards[t-1] += gamma * values[t-1] * time_outs; NULL: off (the reference tasks
   * put no 'time_outs' into their extras, so the switch is inert there) */
  const int64_t* time_outs; float gamma; int32_t pad2_;
  /* normalize_value (cfg/train/*PPO.yaml:47): {mean, std} from lm_rms_export_denorm; values are stored as std * clamp(v, +-5) + mean.
   * NULL: stored as they come */
  const float* value_denorm;
  const LmActionSink* sink;   /* NULL, or: the sampled actions also go straight into the task (see LmActionSink) */
} LmRolloutStep;
typedef struct LmRolloutBuffers {
  float* obses;    /* [T,N,64] */
  float* actions;  /* [T,N,12] */
  float* mus;      /* [T,N,12] */
  float* values;   /* [T,N]    */
  float* rewards;  /* [T,N]  rew * reward_scale (rl_games rewards_shaper scale_value) */
  float* dones;    /* [T,N]  fp32 0/1 */
  float* neglogp;  /* [T,N]  0.5 sum(((a - mu)/sigma)^2) + 0.5 log(2 pi) 12 + sum(log_std)  (= -log_prob of skrl) */
  int32_t horizon; int32_t num_envs;
  float reward_scale;
} LmRolloutBuffers;
typedef struct LmRollout LmRollout;
LmRollout* lm_rollout_create(int device);
void lm_rollout_destroy(LmRollout* r);
/* Samples a = noise * sigma + mu (torch.normal's op order), writes slot t of obses / actions / mus / values / dones / neglogp and
 * the shaped reward into slot t-1 (rl_games play_steps stores rewards[n] AFTER env_step(actions[n]): the reward that comes with
 * observation t belongs to the action of slot t-1; nothing at t = 0), optionally copies the actions to `actions_out` [N,12] (the
 * next pre_physics_step input) and, when obs_stats != NULL, folds this step's observation batch into the running mean /
 * variance (batch mean, unbiased batch variance, parallel-variance combine).  One launch.
 * lm_rollout_finish closes the rollout with the step that follows slot T-1: reward into slot T-1, last_values / last_dones [N]
 * (the bootstrap inputs of lm_gae). */
int lm_rollout_write(LmRollout* r, const LmRolloutStep* step, const LmRolloutBuffers* buf, int32_t t, float* actions_out,
                     const LmRunningStats* obs_stats, void* stream);
int lm_rollout_finish(LmRollout* r, const LmRolloutStep* step, const LmRolloutBuffers* buf, float* last_values, float* last_dones,
                      void* stream);
/* Folds a batch x [n,dim] into running statistics (value / return normalisation, or observations outside a rollout). */
int lm_rms_update(LmRollout* r, const float* x, int64_t n, const LmRunningStats* stats, void* stream);
/* fp32 vectors for lm_mlp_forward(norm_mean, norm_inv_std): mean and the variant's 1/std form. */
int lm_rms_export(const LmRunningStats* stats, float* mean_f32, float* inv_std_f32, void* stream);
/* {mean, std} of a scalar normaliser for LmRolloutStep.value_denorm. */
int lm_rms_export_denorm(const LmRunningStats* stats, float* mean_std_f32, void* stream);
/* y = clamp((x - mean[0]) * inv_std[0], +-clip) over n floats (rl_games prepare_dataset: values / returns through value_mean_std;
 * clip <= 0: no clamp).  In place allowed. */
int lm_normalize_clamp(const float* x, float* y, int64_t n, const float* mean, const float* inv_std, float clip, void* stream);
/* Advantage normalisation (a - a.mean()) / (a.std() + eps) in two steps so that a sharded job can sum the moments over ranks in
 * between: moments[3] = {sum, sum of squares, count} as doubles. */
int lm_moments(LmRollout* r, const float* x, int64_t n, double* moments, void* stream);
int lm_moment_normalize(const float* x, float* y, int64_t n, const double* moments, float eps, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * PPO update pieces: the loss head and the optimiser step.  Not in the reference repo: rl_games 1.6.0
 * a2c_continuous.calc_gradients / common_losses (cfg/train/JointLocomanipulationPPO.yaml:52-72: e_clip 0.2, clip_value,
 * critic_coef 2, entropy_coef 0, bounds_loss_coef 1e-3, grad_norm 1, truncate_grads) and torch.optim.Adam.
 *
 * lm_ppo_loss: one pass over a minibatch of B samples.  From the fresh network outputs (mu, value, log_std) and the
 * stored rollout columns it produces dLoss/dmu [B,12], dLoss/dvalue [B], dLoss/dlog_std [12] and the scalars rl_games
 * logs, for   loss = mean(a_loss) + 0.5 critic_coef mean(c_loss) - entropy_coef mean(entropy) + bounds_loss_coef mean(b_loss)
 * The caller back-propagates (grad_mu, grad_value) through its network.  stats[8] = {a_loss, c_loss, b_loss, entropy, kl,
 * loss, clip fraction, 0}. */
typedef struct LmPpoBatch {
  const float* mu;          /* [B,12] fresh policy mean                      */
  const float* value;       /* [B]    fresh value                            */
  const float* log_std;     /* [12]   fresh log sigma                        */
  const float* actions;     /* [B,12] rollout actions                        */
  const float* old_neglogp; /* [B]                                           */
  const float* advantages;  /* [B]    (normalised by the caller if it does)  */
  const float* old_values;  /* [B]                                           */
  const float* returns;     /* [B]                                           */
  const float* old_mu;      /* [B,12] for the KL estimate; may be NULL       */
  const float* old_log_std; /* [12]   for the KL estimate; may be NULL       */
  int64_t batch;
  float e_clip; float critic_coef; float entropy_coef; float bounds_loss_coef;
  int32_t clip_value;       /* rl_games clip_value                           */
  int32_t update_old;       /* != 0: after use, old_mu <- mu and old_log_std <- log_std in place (rl_games dataset.update_mu_sigma) */
} LmPpoBatch;
int lm_ppo_loss(LmRollout* workspace, const LmPpoBatch* batch, float* grad_mu, float* grad_value, float* grad_log_std,
                float* stats, void* stream);
/* Weight gradients of the MLP actor-critic on tcgen05.  Activations and pre-activation gradients of a minibatch
 * travel as bf16 "tile blobs": a [B, D] matrix stored per 128-sample tile in the UMMA canonical K-major operand layout
 * (lm_mlp_blob_bytes / _pack / _unpack).  lm_mlp_wgrad accumulates, in TMEM over the whole minibatch,
 *   dW1 = dZ1^T A0, dW2 = dZ2^T A1, dW3 = dZ3^T A2, dW_head = dH^T A3 and the bias gradients (column sums)
 * by re-reading the same blobs as MN-major operands (no transpose pass), and writes them in nn.Linear layouts.
 * dH [B,16]: columns 0..11 dLoss/dmu, 12 dLoss/dvalue, 13..15 zero.  workspace: workspace_ctas * lm_mlp_wgrad_workspace_floats()
 * floats (one partial per CTA, summed in CTA order: deterministic). */
/* Training forward: lm_mlp_forward that also leaves the bf16 operands of every layer (A0 = normalised observations [n,64],
 * A1 [n,256], A2 [n,128], A3 [n,64]) as tile blobs.  lm_mlp_backward: from dLoss/dmu [n,12] and dLoss/dvalue [n] (lm_ppo_loss) and
 * the saved activations to the pre-activation gradients dH [n,16], dZ3 [n,64], dZ2 [n,128], dZ1 [n,256] as tile blobs (three
 * GEMMs against the transposed weights resident in shared memory, ELU' from the saved activation). */
int lm_mlp_forward_train(LmMlp* mlp, const float* obs, const float* norm_mean, const float* norm_inv_std, float norm_clip, float* mu,
                         float* value, void* a0, void* a1, void* a2, void* a3, int32_t n, void* stream);
int lm_mlp_backward(LmMlp* mlp, const float* grad_mu, const float* grad_value, const void* a1, const void* a2, const void* a3, void* dh,
                    void* dz3, void* dz2, void* dz1, int32_t n, void* stream);
int64_t lm_mlp_blob_bytes(int64_t rows, int32_t d);
int lm_mlp_blob_pack(const float* x, void* blob, int64_t rows, int32_t d, void* stream);
int lm_mlp_blob_unpack(const void* blob, float* x, int64_t rows, int32_t d, void* stream);
int32_t lm_mlp_wgrad_workspace_floats(void);
int lm_mlp_wgrad(const void* dz1, const void* dz2, const void* dz3, const void* dh, const void* a0, const void* a1, const void* a2,
                 const void* a3, int64_t rows, float* workspace, int32_t workspace_ctas, float* g_w1, float* g_b1, float* g_w2,
                 float* g_b2, float* g_w3, float* g_b3, float* g_w_mu, float* g_b_mu, float* g_w_value, float* g_b_value, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * One-shot all-reduce over NVLink peer memory (the PPO gradient all-reduce and the episode-stat reduction — the
 * only collectives of the path; the reference only forwards rl_games' multi_gpu flag, cfg/config.yaml:29,
 * scripts/rlgames_train.py:93-96).  One process per GPU.  lm_comm_create allocates this rank's receive memory;
 * lm_comm_get_handle exports it as a cudaIpc handle (lm_comm_handle_bytes() bytes) which the caller exchanges between the
 * ranks through its own channel (e.g. torch.distributed.all_gather); lm_comm_open maps the peers.  lm_comm_allreduce sums
 * n <= max_floats floats in place over the ranks: every rank pushes its values into every peer's slot, raises a flag and
 * sums the slots in rank order — the result is bit-identical on every rank.  Every rank must call the collectives of one
 * communicator in the same order.  CUDA-graph capturable (the epoch counter lives on the device).
 * lm_mlp_wgrad_allreduce = lm_mlp_wgrad whose final reduction pushes each gradient element to the peers as its partial sum is
 * formed (compute and collective in ONE kernel).  `arena` [arena_floats] is the flat buffer every gradient tensor is a view
 * of; two extra ranges of it (log_std gradient, loss-head statistics) ride in the same message; on return the whole arena
 * holds the sum over ranks.  lm_comm_status: 0, or 1 + r after a time-out waiting for rank r (synchronises the device). */
typedef struct LmComm LmComm;
int lm_comm_handle_bytes(void);
LmComm* lm_comm_create(int device, int rank, int world, int64_t max_floats);
int lm_comm_get_handle(LmComm* c, void* handle_out);
int lm_comm_open(LmComm* c, const void* handles);
void lm_comm_destroy(LmComm* c);
int lm_comm_allreduce(LmComm* c, float* data, int64_t n, void* stream);
/* The same, with elements [gather_off, gather_off + n_f32 + n_i64 + 1) of the message taken from elsewhere instead of `data`: n_f32
 * fp32 values times f32_scale, then n_i64 int64 counters converted to fp32, then the constant one_more (this rank's env count):
 * the episode statistics join the gradient arena's tail in the same launch. */
int lm_comm_allreduce_gather(LmComm* c, float* data, int64_t n, int64_t gather_off, const float* f32_src, int32_t n_f32,
                             float f32_scale, const int64_t* i64_src, int32_t n_i64, float one_more, void* stream);
int lm_comm_status(LmComm* c);
int lm_mlp_wgrad_allreduce(LmComm* c, const void* dz1, const void* dz2, const void* dz3, const void* dh, const void* a0, const void* a1,
                           const void* a2, const void* a3, int64_t rows, float* workspace, int32_t workspace_ctas, float* g_w1,
                           float* g_b1, float* g_w2, float* g_b2, float* g_w3, float* g_b3, float* g_w_mu, float* g_b_mu,
                           float* g_w_value, float* g_b_value, float* arena, int64_t arena_floats, int64_t extra_off0,
                           int64_t extra_n0, int64_t extra_off1, int64_t extra_n1, void* stream);

/* torch.nn.utils.clip_grad_norm_(max_norm) (skipped when max_norm <= 0) followed by one torch.optim.Adam step over a flat
 * fp32 parameter arena (the layout dist.FusedHorizonReducer keeps): grads are first multiplied by grad_scale
 * (1 / world size after the all-reduce sum).  `step` counts from 1.  One launch; n <= 2^24 parameters. */
int lm_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                 float beta2, float eps, float weight_decay, int32_t step, float max_grad_norm, float grad_scale,
                 float* grad_norm_out, void* stream);
/* The same step with the learning rate kept on the device and rl_games' adaptive schedule (lr_schedule: adaptive, legacy = per
 * minibatch; cfg/train/JointLocomanipulationPPO.yaml:57-58) applied inside the launch: the step runs with *lr, then
 *   kl = *kl * kl_scale;  kl > 2 thr: lr = max(lr / 1.5, lr_min);  kl < thr / 2: lr = min(lr * 1.5, lr_max)
 * for the next one.  `kl` usually points into the all-reduced gradient arena's tail (sum of the ranks' minibatch KL, kl_scale =
 * 1 / world size: every rank takes the same decision).  kl == NULL: *lr is used and left alone.  sched == NULL: `lr` argument. */
typedef struct LmAdamSchedule {
  float* lr;            /* [1] device */
  const float* kl;      /* [1] device, may be NULL */
  float kl_scale; float kl_threshold; float lr_min; float lr_max;
} LmAdamSchedule;
int lm_adam_step_scheduled(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const LmAdamSchedule* sched,
                           float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step, float max_grad_norm,
                           float grad_scale, float* grad_norm_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Observation / action noise: utils/domain_randomization/randomize.py:212-305, applied inside
 * VecEnvRLGames.step (envs/vec_env_rlgames.py:57-58,70-72).  Per env: a step counter, zeroed on reset; "on_reset" noise is
 * redrawn for the reset envs, kept in `correlated` [N,dim] and applied to every env every step; "on_interval" noise is
 * drawn afresh for the envs whose counter reached frequency_interval (which zeroes it).  noise = scale * draw + shift
 * (exp() of it for loguniform): gaussian draw ~ N(0,1), scale = std, shift = mean; uniform draw ~ U[0,1), scale = hi - lo,
 * shift = lo; loguniform scale = log hi - log lo, shift = log lo. */
enum { LM_NOISE_ADDITIVE = 0, LM_NOISE_SCALING = 1 };
enum { LM_DIST_GAUSSIAN = 0, LM_DIST_UNIFORM = 1, LM_DIST_LOGUNIFORM = 2 };
typedef struct LmNoiseSpec {
  int32_t enabled; int32_t operation; int32_t distribution; int32_t frequency_interval;   /* frequency_interval: on_interval only */
  float scale; float shift;
} LmNoiseSpec;
typedef struct LmNoiseState {
  int32_t* counter;      /* [N]      _observations_counter_buffer / _actions_counter_buffer (randomize.py:185,210) */
  float* correlated;     /* [N,dim]  _observations_correlated_noise / _actions_correlated_noise (:186,211)          */
  int32_t num_envs; int32_t dim;   /* dim: a multiple of 4 (64 observations, 12 actions) */
  LmNoiseSpec on_reset; LmNoiseSpec on_interval;
} LmNoiseState;
typedef struct LmNoise LmNoise;
LmNoise* lm_noise_create(int device);
void lm_noise_destroy(LmNoise* h);
/* In place on buffer [N,dim].  draws_reset / draws_interval: rows of N(0,1) (gaussian) or U[0,1) draws from the caller's
 * generator, or NULL for the kernel's own counter-based generator (Philox4x32-10 keyed by seed, stream position `step`
 * and the element index: no host synchronisation, no draw tensors in HBM).  compact_draws != 0: row j of a draw tensor
 * belongs to the j-th affected env in ascending order, as the reference draws them (size (k, dim), randomize.py:265,291);
 * otherwise row e belongs to env e. */
int lm_noise_apply(LmNoise* h, const LmNoiseState* state, float* buffer, const int64_t* reset_buf, const float* draws_reset,
                   const float* draws_interval, int32_t compact_draws, uint64_t seed, uint64_t step, void* stream);

/* End-to-end host-buffer entry used by bench.py's e2e leg: one env step with inputs in pinned
 * host memory (state of both robots packed as the LmRobotViews fields, actions, uniforms) and
 * outputs (obs, rew, reset) read back.  Copies run on `stream`; returns after synchronising it. */
typedef struct LmHostStep {
  const float* actions_host;   /* [N,12] */
  const float* uniforms_host;  /* [N,3]  */
  const float* loco_state_host; int64_t loco_state_floats; float* loco_state_dev;
  const float* mani_state_host; int64_t mani_state_floats; float* mani_state_dev;
  float* actions_dev; float* uniforms_dev;
  float* obs_host; float* rew_host; int64_t* reset_host;
  /* Optional dedicated host->device copy stream (NULL: the copies run on `stream`).  When several env shards hand the SAME
   * copy stream to their calls, their input copies queue back to back at full PCIe bandwidth (instead of sharing it between
   * concurrently running copies), and each shard's kernels start on the shard's own stream as soon as that shard's inputs have
   * landed: shard k's kernels and device->host copies run under shard k+1's input copy.  The library orders the two streams
   * with events it owns (inputs landed -> kernels; kernels done with the state arena -> next step's copy). */
  void* h2d_stream;
} LmHostStep;
/* The policy forward that follows the env step in the same call (lm_task_step_host_policy_async): kind 0 none, 1 MLP actor-critic
 * (handle = LmMlp*), 2 GNN on tcgen05 (LmGnn*, lm_gnn_forward_bf16), 3 GNN fp32 (lm_gnn_forward).  mu_host / value_host
 * (pinned, may be NULL) receive the outputs by device->host copies on `stream`. */
typedef struct LmHostPolicy {
  int32_t kind; int32_t pad_;
  void* handle;
  const float* norm_mean; const float* norm_inv_std; float norm_clip; int32_t pad2_;   /* MLP only; both NULL: no standardisation */
  float* mu_dev; float* value_dev;       /* [N,12], [N] */
  float* mu_host; float* value_host;
} LmHostPolicy;
int lm_task_step_host(LmTask* task, const LmRobotViews* loco, const LmRobotViews* mani,
                      const float* env_pos, const LmTaskBuffers* buf, const LmHostStep* io, void* stream);
/* Same, but returns as soon as the copies and kernels are enqueued on `stream` (no host synchronisation).  A batch
 * split into env shards (one task + one stream per shard, the layout multi-GPU sharding uses) steps all shards
 * back to back so that shard k's host->device copies overlap shard k-1's kernels and device->host copies (PCIe is
 * full duplex); the caller then waits with lm_stream_synchronize on each shard's stream. */
int lm_task_step_host_async(LmTask* task, const LmRobotViews* loco, const LmRobotViews* mani,
                            const float* env_pos, const LmTaskBuffers* buf, const LmHostStep* io, void* stream);
/* lm_task_step_host_async followed by the policy forward on the step's observations and the device->host copies of its outputs,
 * all enqueued by ONE call (one shard of a host-driven step costs one library call instead of five). */
int lm_task_step_host_policy_async(LmTask* task, const LmRobotViews* loco, const LmRobotViews* mani, const float* env_pos,
                                   const LmTaskBuffers* buf, const LmHostStep* io, const LmHostPolicy* policy, void* stream);
int lm_stream_synchronize(void* stream);
/* Page-locked host memory for the host-buffer entries (cudaHostAlloc).  write_combined != 0: cudaHostAllocWriteCombined — for
 * buffers the host only WRITES (the physics state on its way to the device): PCIe reads then bypass the CPU caches. */
void* lm_host_alloc(int64_t bytes, int32_t write_combined);
void lm_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* LM_B200_H */
