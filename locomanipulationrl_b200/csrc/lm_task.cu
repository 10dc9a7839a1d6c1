// Task hot path for sm_100a: reset rank scan, pre_physics_step (+reset_idx scatter, goal resample,
// action -> target) and post_physics_step (state read-back, observation assembly, rot-distance
// reward, termination, success bookkeeping) — one env per thread, one 128-env tile per CTA, every
// per-env array staged through shared memory by 1-D TMA bulk copies (the rows of a tile are one
// contiguous span of the row-major [N,k] tensors the reference's API fixes).
//
// Reference restated (paths under /root/reference/RobotLearning/omniisaacgymenvs/):
//   tasks/quadruped_pose_control_tasks/quadruped_pose_control.py:200-633        (QPC)
//   tasks/quadruped_manipulate_plate/quadruped_manipulate_plate.py:213-652      (QMP)
//   tasks/joint_train_locomanipulation/joint_locomanipulation_vertical.py:248-802 (JLV)
//   robot/base/robot.py:276-321,444-461; objects/base/rigid_object.py:36-58; utils/math.py:7-193
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "lm_common.cuh"
#include "lm_host.h"

namespace lm {

constexpr int T = 128;            // envs per tile == threads per CTA
constexpr int OBS_PAD = 68;       // padded obs row (floats): 272 B stride -> conflict-free 128-bit smem access
constexpr int STATS = 16;         // per-CTA partials: 7 reward terms, nan, below -50, goal_reset, reset, pad

struct Derived {                  // host-computed constants (lm_task_create)
  float R0[9];                    // rot_matrices_from_quat(mani default quaternion)  (QMP:149, math.py:7-19)
  float p0inv[3];                 // bmm(R0^T, -p0)                                   (math.py:165-166)
  float corners_world[4][3];      // transform_vectors(q0, p0, corners)               (QMP:162-165)
  int32_t half;                   // envs per half (N for single-task kinds)
  int32_t tiles_per_half;
  int32_t num_tiles;
  int32_t dof_identity;           // dof_index == 0..11
  int32_t scan_per;               // tiles per scan block: reset rank base of tile t = scan_base[t / scan_per] + tile_offsets[t]
};

struct KernelParams {
  LmTaskParams p;
  Derived d;
};

// ------------------------------------------------------------------------------------------------
// tile geometry: tiles never straddle the locomotion / manipulation boundary
// ------------------------------------------------------------------------------------------------
struct Tile {
  int is_mani;      // which robot's views this tile reads
  int hr0;          // first row in the half-arrays
  int e0;           // first global env id
  int n;            // valid envs in this tile
};
__device__ __forceinline__ Tile tile_of(const KernelParams& kp, int t) {
  Tile tl;
  int second = t >= kp.d.tiles_per_half && kp.p.kind == LM_KIND_COTRAIN;
  int lt = second ? t - kp.d.tiles_per_half : t;
  tl.is_mani = (kp.p.kind == LM_KIND_MANI) || second;
  tl.hr0 = lt * T;
  tl.e0 = (second ? kp.d.half : 0) + tl.hr0;
  tl.n = min(T, kp.d.half - tl.hr0);
  return tl;
}

// staged load of `nfloats` contiguous floats: TMA when the span is a multiple of 16 B, else cooperative.
struct Stager {
  uint64_t* bar;
  uint32_t bytes;
  bool use_tma;
  __device__ __forceinline__ void load(float* dst, const float* src, int nfloats) {
    if (use_tma) {
      if (threadIdx.x == 0) {
        tma_load_1d(dst, src, (uint32_t)nfloats * 4u, bar);
        bytes += (uint32_t)nfloats * 4u;
      }
    } else {
      for (int i = threadIdx.x; i < nfloats; i += blockDim.x) dst[i] = __ldg(src + i);
    }
  }
};

// ------------------------------------------------------------------------------------------------
// K_scan: per-tile reset counts + exclusive offsets (the reset rank = goal index of QPC:206 / math.py:184)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(T) k_reset_count(const KernelParams kp, const int64_t* __restrict__ reset_buf,
                                                   int32_t* __restrict__ tile_counts, int32_t* __restrict__ tile_offsets,
                                                   int32_t* __restrict__ scan_base, int32_t* __restrict__ num_reset,
                                                   unsigned int* __restrict__ ticket) {
  __shared__ int warp_counts[T / 32];
  __shared__ int is_last;
  __shared__ int scan_s[T];
  const Tile tl = tile_of(kp, blockIdx.x);
  const int r = threadIdx.x;
  int flag = 0;
  if (r < tl.n) flag = reset_buf[tl.e0 + r] != 0;
  unsigned b = __ballot_sync(0xffffffffu, flag);
  if ((r & 31) == 0) warp_counts[r >> 5] = __popc(b);
  __syncthreads();
  if (r == 0) {
    int c = 0;
#pragma unroll
    for (int w = 0; w < T / 32; ++w) c += warp_counts[w];
    tile_counts[blockIdx.x] = c;
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // last CTA: exclusive scan over all tile counts, chunked per thread (absolute offsets: scan_base stays zero)
  const int nt = kp.d.num_tiles;
  for (int i = r; i < (nt + kp.d.scan_per - 1) / kp.d.scan_per; i += T) scan_base[i] = 0;
  const int per = (nt + T - 1) / T;
  int lo = r * per, hi = min(nt, lo + per), s = 0;
  for (int i = lo; i < hi; ++i) s += __ldcg(tile_counts + i);
  scan_s[r] = s;
  __syncthreads();
  if (r == 0) {
    int acc = 0;
    for (int i = 0; i < T; ++i) { int v = scan_s[i]; scan_s[i] = acc; acc += v; }
    num_reset[0] = acc;
    if (kp.p.kind == LM_KIND_LOCO) num_reset[1] = acc;
    if (kp.p.kind == LM_KIND_MANI) num_reset[1] = 0;
    *ticket = 0;
  }
  __syncthreads();
  int acc = scan_s[r];
  for (int i = lo; i < hi; ++i) {
    if (kp.p.kind == LM_KIND_COTRAIN && i == kp.d.tiles_per_half) num_reset[1] = acc;   // resets in the locomotion half
    tile_offsets[i] = acc;
    acc += __ldcg(tile_counts + i);
  }
}

// ------------------------------------------------------------------------------------------------
// K_pre: clamp + stash actions, action -> target, reset_idx scatter + goal resample
// ------------------------------------------------------------------------------------------------
struct PreArgs {
  LmRobotViews loco, mani;
  const float* env_pos;
  LmTaskBuffers buf;
  const float* actions;
  const float* uniforms;
  int32_t* tile_offsets;
  int32_t* scan_base;
  int parity;                      // actions_ring slot that receives current_actions
  LmPdState pd;                    // *_position_control variants (joint_pos == NULL otherwise)
  // deferred finish (lm_task_set_defer_finish): the post kernel only published per-tile partial sums and reset counts.  Every CTA
  // then derives its reset-rank base from the dense per-tile counts itself (`tile_counts`, one 128-bit load per thread up to 512
  // tiles), and ONE extra CTA (blockIdx.x == num_tiles) runs the finish pass of the previous post step — extras, counters — next to
  // the reset bodies instead of as the serial tail of the post kernel.
  const int32_t* tile_counts;      // non-NULL: rank base = sum of tile_counts[0 .. blockIdx.x)
  float* partials;                 // non-NULL: the extra CTA runs the finish pass over these
  unsigned int* ticket;
};

__device__ __forceinline__ void store3(float* p, float a, float b, float c) { p[0] = a; p[1] = b; p[2] = c; }

// reset_idx for one env (QPC:230-299 / QMP:238-309 / JLV:272-377): defaults into the PhysX views, goal resample from
// uniform row `rank`, history buffers cleared.  `role` splits the work over two lanes: 0 = goal quaternion + pose
// indicator (the long double-precision chain), 1 = joint / root / plate defaults and history; -1 = both.
__device__ __forceinline__ void reset_one_env(const KernelParams& kp, const PreArgs& a, const LmRobotViews& V, float* ring_last,
                                              int is_mani, int e, int hr, int rank, int role) {
  const LmTaskParams& P = kp.p;
  const float ox = a.env_pos[hr * 3 + 0], oy = a.env_pos[hr * 3 + 1], oz = a.env_pos[hr * 3 + 2];
  if (role != 1) {
    const float* u = a.uniforms + (size_t)rank * 3;
    const Quat g = goal_from_uniforms(P, u[0], u[1], u[2]);
    float* gq = a.buf.goal_quaternions + (size_t)e * 4;
    gq[0] = g.w; gq[1] = g.x; gq[2] = g.y; gq[3] = g.z;
    a.buf.reset_env_ids[rank] = e;
    float* iq = V.indicator_quat + hr * 4;
    if (!is_mani) {
      iq[0] = g.w; iq[1] = -g.x; iq[2] = -g.y; iq[3] = -g.z;                             // conj(goal), QPC:282
      store3(V.indicator_pos + hr * 3, add(ox, P.indicator_loco_position[0]), add(oy, P.indicator_loco_position[1]),
             add(oz, P.indicator_loco_position[2]));                                    // rigid_object.py:45
    } else {
      float Rg[9], M[9], R0[9];                                                         // rotate_orientations(q0, goal), QMP:291
#pragma unroll
      for (int i = 0; i < 9; ++i) R0[i] = kp.d.R0[i];
      rot_from_quat_scipy(g, Rg);
      mat3_mul<false>(R0, Rg, M);
      const Quat iqv = quat_from_rot_scipy(M);
      iq[0] = iqv.w; iq[1] = iqv.x; iq[2] = iqv.y; iq[3] = iqv.z;
      store3(V.indicator_pos + hr * 3, add(ox, P.indicator_mani_position[0]), add(oy, P.indicator_mani_position[1]),
             add(oz, P.indicator_mani_position[2]));
    }
  }
  if (role != 0) {
    float4* jp = reinterpret_cast<float4*>(V.dof_pos + (size_t)hr * LM_NUM_DOF);
    float4* jv = reinterpret_cast<float4*>(V.dof_vel + (size_t)hr * LM_NUM_DOF);
#pragma unroll
    for (int k = 0; k < LM_NUM_DOF / 4; ++k) {
      jp[k] = is_mani ? make_float4(P.mani_init_joint_pos[4 * k], P.mani_init_joint_pos[4 * k + 1], P.mani_init_joint_pos[4 * k + 2], P.mani_init_joint_pos[4 * k + 3])
                      : make_float4(P.loco_init_joint_pos[4 * k], P.loco_init_joint_pos[4 * k + 1], P.loco_init_joint_pos[4 * k + 2], P.loco_init_joint_pos[4 * k + 3]);
      jv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (!is_mani) {
      store3(V.root_pos + hr * 3, add(P.loco_default_position[0], ox), add(P.loco_default_position[1], oy),
             add(P.loco_default_position[2], oz));                                      // robot.py:420
      *reinterpret_cast<float4*>(V.root_quat + hr * 4) = make_float4(P.loco_default_quaternion[0], P.loco_default_quaternion[1],
                                                                    P.loco_default_quaternion[2], P.loco_default_quaternion[3]);
      store3(V.root_lin_vel + hr * 3, 0.f, 0.f, 0.f);
      store3(V.root_ang_vel + hr * 3, 0.f, 0.f, 0.f);
    } else {
      store3(V.plate_pos + hr * 3, add(ox, P.obj_default_position[0]), add(oy, P.obj_default_position[1]),
             add(oz, P.obj_default_position[2]));
      *reinterpret_cast<float4*>(V.plate_quat + hr * 4) = make_float4(P.mani_default_quaternion[0], P.mani_default_quaternion[1],
                                                                     P.mani_default_quaternion[2], P.mani_default_quaternion[3]);
      store3(V.plate_lin_vel + hr * 3, 0.f, 0.f, 0.f);
      store3(V.plate_ang_vel + hr * 3, 0.f, 0.f, 0.f);
    }
    float4* ljv = reinterpret_cast<float4*>(a.buf.last_joint_vel + (size_t)e * LM_NUM_ACT);
    float4* la = reinterpret_cast<float4*>(ring_last + (size_t)e * LM_NUM_ACT);
    float4* lt = reinterpret_cast<float4*>(a.buf.last_base_tip + (size_t)e * 12);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ljv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      la[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      lt[k] = make_float4(P.default_base_tip[(4 * k) / 3][(4 * k) % 3], P.default_base_tip[(4 * k + 1) / 3][(4 * k + 1) % 3],
                          P.default_base_tip[(4 * k + 2) / 3][(4 * k + 2) % 3], P.default_base_tip[(4 * k + 3) / 3][(4 * k + 3) % 3]);
    }
    a.buf.reset_buf[e] = 0; a.buf.progress_buf[e] = 0; a.buf.successes[e] = 0;
    a.buf.consecutive_successes[e] = 0; a.buf.goal_reset_buf[e] = 0;
    if (a.pd.joint_pos != nullptr) {                                                    // …position_control.py:377-386
      const float* init = is_mani ? P.mani_init_joint_pos : P.loco_init_joint_pos;
      float4* jq = reinterpret_cast<float4*>(a.pd.joint_pos + (size_t)e * LM_NUM_ACT);
      float4* jd = reinterpret_cast<float4*>(a.pd.joint_vel + (size_t)e * LM_NUM_ACT);
      float4* tg = reinterpret_cast<float4*>(a.pd.targets + (size_t)e * LM_NUM_ACT);
      float4* ltg = reinterpret_cast<float4*>(a.pd.last_targets + (size_t)e * LM_NUM_ACT);
      float4* se = reinterpret_cast<float4*>(a.pd.targets_se + (size_t)e * LM_NUM_ACT);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float4 d = make_float4(init[4 * k], init[4 * k + 1], init[4 * k + 2], init[4 * k + 3]);
        jq[k] = d; tg[k] = d; ltg[k] = d;
        jd[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        se[k] = make_float4(a.pd.init_se[4 * k], a.pd.init_se[4 * k + 1], a.pd.init_se[4 * k + 2], a.pd.init_se[4 * k + 3]);
      }
    }
  }
}

// direct reset_idx(env_ids) call (random_policy.py-style drivers): env k of the list consumes uniform row k
__global__ void k_reset_idx(const KernelParams kp, const PreArgs a, const int64_t* __restrict__ env_ids, int count) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int e = (int)env_ids[k];
  if (e < 0 || e >= kp.p.num_envs) return;
  const int second = (kp.p.kind == LM_KIND_COTRAIN) && e >= kp.d.half;
  const int is_mani = (kp.p.kind == LM_KIND_MANI) || second;
  const int hr = second ? e - kp.d.half : e;
  float* ring_last = a.buf.actions_ring + (size_t)(a.parity ^ 1) * kp.p.num_envs * LM_NUM_ACT;
  reset_one_env(kp, a, is_mani ? a.mani : a.loco, ring_last, is_mani, e, hr, k, -1);
}

// ------------------------------------------------------------------------------------------------
// K_post: per-env math shared by both tilings
// ------------------------------------------------------------------------------------------------
struct PostArgs {
  LmRobotViews loco, mani;
  const float* env_pos;
  LmTaskBuffers buf;
  float* partials;                 // [num_tiles, STATS] fp32 per-CTA partial sums
  unsigned int* ticket;
  int32_t* tile_offsets;           // by-product: exclusive scan of per-tile reset counts for the next pre_step,
  int32_t* scan_base;              //   as scan_base[t / scan_per] + tile_offsets[t]
  int32_t* tile_counts;            // dense per-tile reset counts, always written (k_pre's own scan after a deferred finish)
  int inline_finish;               // the last CTA runs the finish pass (else k_post_finish is launched behind the kernel)
  const float* pd_targets;         // *_position_control variants: obs[40:64] = scale * targets, scale * last_targets
  const float* pd_last_targets;    //   (…position_control.py:438-452); NULL otherwise
  float pd_target_obs_scale;
  float* pd_joint_pos;             // single-task *_position_control: the observed joint state becomes the PD loop's carried
  float* pd_joint_vel;             //   state (…position_control.py:392-393); NULL otherwise
  int parity;
  int clamp_observations;          // store clamp(obs, ±clip_observations) (fused _process_data)
  long long* timeline;             // diagnostic: [num_tiles, 8] globaltimer stamps (NULL in production)
};

__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// (compiled into the kStamp instantiation of k_post_lean only, which is launched while a timeline buffer is registered)
template <bool kStamp>
__device__ __forceinline__ void stamp(const PostArgs& a, int k) {
  if (kStamp && a.timeline != nullptr && threadIdx.x == 0) a.timeline[(size_t)blockIdx.x * 8 + k] = globaltimer_ns();
}
__device__ __forceinline__ void stamp_at(const PostArgs& a, int k) {   // caller picks the thread
  if (a.timeline != nullptr) a.timeline[(size_t)blockIdx.x * 8 + k] = globaltimer_ns();
}

struct EnvIn {
  float jp[12], jv[12], last_jv[12], cur[12], last[12];
  Vec3 O, Pw, lin, ang;            // env origin, world position of base | plate, velocities
  Quat q, goal;                    // orientation of base | plate, goal
  int progress, reset, succ_prev, consec;   // int64 at the boundary; values are small
};
struct EnvOut {
  Quat cur_q;                      // ground | plate quaternion in the robot frame (states row)
  float rew;
  int progress, consec;
  int reset, success, goal_reset;
  float stats[STATS];
};

// Emits one 16-byte chunk k (0..15) of the 64-float observation row.
struct ObsToGlobal {
  float4* row; float clip; int clamp; bool on = true;
  __device__ __forceinline__ void operator()(int k, float a, float b, float c, float d) const {
    if (!on) return;
    if (clamp) {
      a = fminf(fmaxf(a, -clip), clip); b = fminf(fmaxf(b, -clip), clip);
      c = fminf(fmaxf(c, -clip), clip); d = fminf(fmaxf(d, -clip), clip);
    }
    row[k] = make_float4(a, b, c, d);
  }
};

// The per-env math is split into a POSE half (orientation / frame quantities, obs[0:16], orientation reward, success
// state machine, termination) and a LIMB half (joint sums and limits, obs[16:64], tips / corners / knees).  The single-
// kernels run both in one thread, in the order that suits their staging: k_post_lean needs the limb half to consume the
// staged knees before observation chunks are emitted.  FrameShare (pose -> limb) and LimbOut (limb -> pose) are the
// only values crossing between the halves.
struct FrameShare {                // pose -> limb
  float R[9];                      // rot_matrices_from_quat(q)   (math.py:7-19)
  float pinv[3];                   // bmm(R^T, -pl)               (math.py:165-166)
  float plz;                       // local z of the base
};
struct LimbOut {                   // limb -> pose
  float jacc, arate;               // joint-acceleration / action-rate reward terms
  int brk, rst, corner_low, knee_low;
};

// ---- limb half, part 1 (no frame needed): acceleration / action-rate sums (robot.py:290; QPC:443,446), joint-limit
//      flags (QPC:466-504 / JLV:574-597) and obs[16:64] (QPC:367-370)
//      tail_cur / tail_last (global rows, NULL unless a *_position_control task): obs[40:64] = scale * targets,
//      scale * last_targets instead of the actions (…position_control.py:438-452)
template <typename Sink>
__device__ __forceinline__ void limb_joints(const KernelParams& kp, const float jp[12], const float jv[12], const float last_jv[12],
                                            float cur[12], float last[12], const Sink& sink, LimbOut& lo,
                                            const float* tail_cur = nullptr, const float* tail_last = nullptr, float tail_scale = 1.0f) {
  const LmTaskParams& P = kp.p;
  float acc_sum = 0.0f, rate_sum = 0.0f;
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    const float acc = fdiv(sub(jv[j], last_jv[j]), P.time_interval);
    acc_sum = add(acc_sum, mul(fabsf(acc), P.joint_acc_scale));
    rate_sum = add(rate_sum, fabsf(sub(last[j], cur[j])));
  }
  lo.jacc = acc_sum;
  lo.arate = mul(rate_sum, P.action_rate_scale);
  int brk = 0, rst = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float d = fabsf(sub(jp[5 + 2 * k], jp[4 + 2 * k]));
    brk |= (d < P.min_joint_23_diff) | (d > P.max_joint_23_diff);
    rst |= (d < P.reset_min_joint_23_diff) | (d > P.reset_max_joint_23_diff);
    const bool mirrored = P.joint1_mirrored && (k == 1 || k == 2);
    const float lo1 = mirrored ? -P.max_joint_1_pos : P.min_joint_1_pos, hi1 = mirrored ? -P.min_joint_1_pos : P.max_joint_1_pos;
    const float rlo = mirrored ? -P.reset_max_joint_1_pos : P.reset_min_joint_1_pos;
    const float rhi = mirrored ? -P.reset_min_joint_1_pos : P.reset_max_joint_1_pos;
    brk |= (jp[k] < lo1) | (jp[k] > hi1);
    rst |= (jp[k] < rlo) | (jp[k] > rhi);
  }
  lo.brk = brk; lo.rst = rst;
  if (tail_cur != nullptr) {       // the action registers are dead: reuse them for the target rows
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float4 c = reinterpret_cast<const float4*>(tail_cur)[k], l = reinterpret_cast<const float4*>(tail_last)[k];
      cur[4 * k] = mul(c.x, tail_scale); cur[4 * k + 1] = mul(c.y, tail_scale); cur[4 * k + 2] = mul(c.z, tail_scale); cur[4 * k + 3] = mul(c.w, tail_scale);
      last[4 * k] = mul(l.x, tail_scale); last[4 * k + 1] = mul(l.y, tail_scale); last[4 * k + 2] = mul(l.z, tail_scale); last[4 * k + 3] = mul(l.w, tail_scale);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    sink(4 + k, mul(P.joint_position_scale, jp[4 * k]), mul(P.joint_position_scale, jp[4 * k + 1]),
         mul(P.joint_position_scale, jp[4 * k + 2]), mul(P.joint_position_scale, jp[4 * k + 3]));
    sink(7 + k, mul(P.joint_velocity_scale, jv[4 * k]), mul(P.joint_velocity_scale, jv[4 * k + 1]),
         mul(P.joint_velocity_scale, jv[4 * k + 2]), mul(P.joint_velocity_scale, jv[4 * k + 3]));
    sink(10 + k, cur[4 * k], cur[4 * k + 1], cur[4 * k + 2], cur[4 * k + 3]);
    sink(13 + k, last[4 * k], last[4 * k + 1], last[4 * k + 2], last[4 * k + 3]);
  }
}

// ---- limb half, part 2: tips into the robot frame (in place), corner / knee height tests
//      tips: 4x3 world positions in, base-frame positions out; knees: 8x3 world positions
__device__ __forceinline__ void limb_frame(const KernelParams& kp, const int is_mani, const Vec3& O, const FrameShare& f, float* tips,
                                           const float* knees, LimbOut& lo) {
  const LmTaskParams& P = kp.p;
  int corner_low = 0, knee_low = 0;
  if (!is_mani) {
    const Vec3 pinv{f.pinv[0], f.pinv[1], f.pinv[2]};
#pragma unroll
    for (int k = 0; k < 4; ++k) {                                                       // QPC:332-335
      const Vec3 t = tf_inverse_apply(f.R, pinv, Vec3{sub(tips[3 * k], O.x), sub(tips[3 * k + 1], O.y), sub(tips[3 * k + 2], O.z)});
      tips[3 * k] = t.x; tips[3 * k + 1] = t.y; tips[3 * k + 2] = t.z;
      const float cz = add(row_dot(f.R + 6, 1, Vec3{P.corners[k][0], P.corners[k][1], P.corners[k][2]}), f.plz);   // QPC:574-579
      corner_low |= cz < P.baseline_corner_height;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) knee_low |= sub(sub(knees[3 * k + 2], O.z), P.baseline_knee_height) <= 0.0f;   // QPC:585-587
  } else {
    const Vec3 p0inv{kp.d.p0inv[0], kp.d.p0inv[1], kp.d.p0inv[2]};
    float R0[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R0[i] = kp.d.R0[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const Vec3 t = tf_inverse_apply(R0, p0inv, Vec3{sub(tips[3 * k], O.x), sub(tips[3 * k + 1], O.y), sub(tips[3 * k + 2], O.z)});   // QMP:343
      tips[3 * k] = t.x; tips[3 * k + 1] = t.y; tips[3 * k + 2] = t.z;
      const float cz = tf_inverse_apply_z(f.R, f.pinv[2], Vec3{kp.d.corners_world[k][0], kp.d.corners_world[k][1], kp.d.corners_world[k][2]});
      corner_low |= cz < P.baseline_corner_height;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {                                                       // plate frame (QMP:576-603)
      const float kz = tf_inverse_apply_z(f.R, f.pinv[2], Vec3{sub(knees[3 * k], O.x), sub(knees[3 * k + 1], O.y), sub(knees[3 * k + 2], O.z)});
      knee_low |= sub(kz, P.baseline_knee_height) <= 0.0f;
    }
  }
  lo.corner_low = corner_low; lo.knee_low = knee_low;
}

struct PoseIn {
  Vec3 O, Pw, lin, ang;
  Quat q, goal;
  int progress, reset, succ_prev, consec;
};
struct PoseMid {                   // pose-half values that wait for the limb half
  float posz, body_height, rot_rew, trans, bonus;
  int goal_reset, success, consec;
};

// ---- pose half, part 1: local pose and the frame the limb half needs
__device__ __forceinline__ Vec3 pose_frame(const PoseIn& in, FrameShare& f) {
  const Vec3 pl{sub(in.Pw.x, in.O.x), sub(in.Pw.y, in.O.y), sub(in.Pw.z, in.O.z)};   // local pose (robot.py:295 / rigid_object.py:54)
  rot_from_quat_scipy(in.q, f.R);
  const Vec3 pinv = tf_inverse_translation(f.R, pl);
  f.pinv[0] = pinv.x; f.pinv[1] = pinv.y; f.pinv[2] = pinv.z;
  f.plz = pl.z;
  return pl;
}

// ---- pose half, part 2: frame quantities per kind, orientation error, obs[0:16], reward terms, success state machine
template <typename Sink>
__device__ __forceinline__ void pose_main(const KernelParams& kp, const int is_mani, const PoseIn& in, const Vec3& pl, const FrameShare& f,
                                          const Sink& sink, Quat& cur_q_out, PoseMid& m) {
  const LmTaskParams& P = kp.p;
  Vec3 pos, lin_o, ang_o;
  Quat cur_q;
  if (!is_mani) {
    pos = quat_rotate_t<true>(in.q, Vec3{-pl.x, -pl.y, -pl.z});                          // QPC:319
    cur_q = quat_conj(in.q);                                                            // QPC:322
    lin_o = quat_rotate_t<true>(in.q, Vec3{-in.lin.x, -in.lin.y, -in.lin.z});           // QPC:326
    ang_o = quat_rotate_t<true>(in.q, Vec3{-in.ang.x, -in.ang.y, -in.ang.z});           // QPC:329
    m.body_height = pl.z;                                                               // QPC:569
  } else {
    const Vec3 p0inv{kp.d.p0inv[0], kp.d.p0inv[1], kp.d.p0inv[2]};
    float R0[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R0[i] = kp.d.R0[i];
    pos = tf_inverse_apply(R0, p0inv, pl);                                              // QMP:326-329
    float M[9];
    mat3_mul<true>(R0, f.R, M);                                                         // QMP:332 (math.py:70-74)
    cur_q = quat_from_rot_scipy(M);
    const Quat q0{P.mani_default_quaternion[0], P.mani_default_quaternion[1], P.mani_default_quaternion[2], P.mani_default_quaternion[3]};
    lin_o = quat_rotate_t<true>(q0, in.lin);                                            // QMP:338
    ang_o = quat_rotate_t<true>(q0, in.ang);
    m.body_height = tf_inverse_apply_z(f.R, f.pinv[2], Vec3{P.mani_default_position[0], P.mani_default_position[1], P.mani_default_position[2]});
  }

  // ---- orientation error (QPC:345-350,430-432)
  Quat diff = quat_mul(cur_q, quat_conj(in.goal));
  const float nrm = norm3_torch(diff.x, diff.y, diff.z);
  const float rot_dist = mul(2.0f, (float)asin((double)fminf(nrm, 1.0f)));
  if (diff.w < 0.0f) { diff.w = -diff.w; diff.x = -diff.x; diff.y = -diff.y; diff.z = -diff.z; }
  const Vec3 up = quat_rotate_t<false>(cur_q, Vec3{0.0f, 0.0f, 1.0f});                  // quat_axis(q, 2)

  // ---- obs[0:16] (QPC:358-366)
  sink(0, mul(P.position_scale, pos.x), mul(P.position_scale, pos.y), mul(P.position_scale, pos.z), up.x);
  sink(1, up.y, up.z, diff.w, diff.x);
  sink(2, diff.y, diff.z, mul(P.linear_vel_scale, lin_o.x), mul(P.linear_vel_scale, lin_o.y));
  sink(3, mul(P.linear_vel_scale, lin_o.z), mul(P.angular_vel_scale, ang_o.x), mul(P.angular_vel_scale, ang_o.y),
       mul(P.angular_vel_scale, ang_o.z));
  cur_q_out = cur_q;

  // ---- reward terms that do not involve the limbs (QPC:432-457)
  m.posz = pos.z;
  m.rot_rew = mul(fdiv(1.0f, add(fabsf(rot_dist), P.rot_eps)), P.quaternion_scale);
  m.trans = mul(norm2_torch(pos.x, pos.y), P.translation_scale);
  m.goal_reset = in.consec > P.max_consecutive_successes;                                // counter BEFORE update (QPC:449)
  m.bonus = mul(P.consecutive_success_reward, (float)m.goal_reset);
  m.success = fabsf(rot_dist) <= P.success_thresh;

  // ---- consecutive-success state machine (QPC:534-545)
  int consec = in.consec;
  const int both = m.success && (in.succ_prev != 0);
  if (!both) consec = 0;
  if (both) consec += 1;
  if (in.succ_prev == 0 && m.success) consec = 1;
  m.consec = consec;
}

// ---- pose half, part 3: total reward (QPC:506-515) and termination (QPC:562-616) once the limb half is known
__device__ __forceinline__ void pose_finish(const KernelParams& kp, const PoseIn& in, const PoseMid& m, const LimbOut& lo, EnvOut& out) {
  const LmTaskParams& P = kp.p;
  const float limit_pen = mul(P.joint_limit_penalty, (float)lo.brk);
  float total = add(add(add(add(add(m.rot_rew, m.trans), lo.jacc), lo.arate), m.bonus), limit_pen);
  int rs = in.reset != 0;
  rs |= m.posz > 0.0f;
  rs |= m.body_height <= P.baseline_height;
  rs |= lo.corner_low | lo.knee_low | lo.rst;
  const float fall = mul(P.fall_penalty, (float)rs);
  total = add(total, fall);
  rs |= m.goal_reset;
  const int progress = in.progress + 1;                                                 // rl_task.py:251
  rs |= progress >= (P.max_episode_length - 1);

  out.rew = total; out.progress = progress; out.consec = m.consec; out.reset = rs; out.success = m.success; out.goal_reset = m.goal_reset;
#pragma unroll
  for (int i = 0; i < STATS; ++i) out.stats[i] = 0.0f;
  out.stats[0] = m.rot_rew; out.stats[1] = m.trans; out.stats[2] = lo.jacc; out.stats[3] = lo.arate; out.stats[4] = m.bonus;
  out.stats[5] = limit_pen; out.stats[6] = fall;
  out.stats[7] = isnan(total) ? 1.0f : 0.0f;
  out.stats[8] = total < -50.0f ? 1.0f : 0.0f;
  out.stats[9] = (float)m.goal_reset;
  out.stats[10] = (float)rs;
}

// Both halves in one thread (k_post: the 93-wide states variant)
template <typename Sink>
__device__ __forceinline__ void env_math(const KernelParams& kp, const int is_mani, const EnvIn& in, float* tips,
                                         const float* knees, const Sink& sink, EnvOut& out) {
  LimbOut lo;
  float cur[12], last[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) { cur[j] = in.cur[j]; last[j] = in.last[j]; }
  limb_joints(kp, in.jp, in.jv, in.last_jv, cur, last, sink, lo);
  PoseIn pin{in.O, in.Pw, in.lin, in.ang, in.q, in.goal, in.progress, in.reset, in.succ_prev, in.consec};
  FrameShare f;
  const Vec3 pl = pose_frame(pin, f);
  limb_frame(kp, is_mani, in.O, f, tips, knees, lo);
  PoseMid m;
  pose_main(kp, is_mani, pin, pl, f, sink, out.cur_q, m);
  pose_finish(kp, pin, m, lo, out);
}

__device__ __forceinline__ void store_flags(const LmTaskBuffers& b, int e, const EnvOut& out) {
  b.rew_buf[e] = out.rew;
  b.progress_buf[e] = (int64_t)out.progress;
  b.reset_buf[e] = (int64_t)out.reset;
  b.successes[e] = (int64_t)out.success;
  b.consecutive_successes[e] = (int64_t)out.consec;
  b.goal_reset_buf[e] = (int64_t)out.goal_reset;
}

// per-CTA partial sums -> partials[tile]; the last CTA turns them into extras / counters / the next reset scan
// (QPC:560,610,619-633 / JLV:758-802).  The last CTA's pass is the serial tail of the launch, so it is kept short:
// thread r owns a contiguous block of tiles (batched 128-bit L2 loads, per-thread sums in double), then one shuffle
// tree / block scan.  Thread->tile mapping and tree shape are fixed, so `extras` are deterministic run to run.
constexpr int FINISH_THREADS = 1024;   // stand-alone finish kernel (large grids)
constexpr int INLINE_FINISH_MAX_TILES = 1024;

struct FinishSmem {
  float red_f[8][STATS];            // per-warp partial sums of this CTA (up to 8 warps)
  double red_d[32][13];             // finish pass: per-warp sums (11 totals + manipulation-half counts 9, 10)
  int warp_scan[32];
  int is_last;
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Finish pass over partials[num_tiles][STATS] by one CTA of kThreads threads.  Thread r owns the contiguous tiles
// [r * scan_per, (r + 1) * scan_per): their partial sums are fetched with batched 128-bit L2 loads and summed in double;
// one shuffle tree gives the totals and one block scan the reset ranks.  Per-tile offsets are written RELATIVE to the
// owner's block and the block bases separately (scan_base), so no second pass over the tiles is needed.
template <int kThreads, bool kSeparate, typename FinArgs>
__device__ __forceinline__ void finish_tail(const KernelParams& kp, const FinArgs& a, FinishSmem& fs) {
  const LmTaskParams& P = kp.p;
  const int r = threadIdx.x, N = P.num_envs, lane = r & 31, warp = r >> 5;
  constexpr int kWarps = kThreads / 32;
  const int nt = kp.d.num_tiles, tph = kp.d.tiles_per_half, per = kp.d.scan_per;
  const int lo = min(nt, r * per), hi = min(nt, lo + per);
  if (kSeparate) {
    // A separately launched pass (deferred finish: an extra CTA of k_pre, or k_post_finish) may be issued twice for one post step —
    // a captured step graph carries the extra CTA, the caller ran lm_task_post_finish at a rollout boundary — and the counters
    // accumulate: only the first one after a post step acts.
    if (r == 0) fs.is_last = (int)__ldcg(a.ticket + 1);
    __syncthreads();
    if (fs.is_last == 0) return;
  }
  double acc[13];
#pragma unroll
  for (int i = 0; i < 13; ++i) acc[i] = 0.0;
  int64_t cnt[6] = {0, 0, 0, 0, 0, 0};                // success / reset counters: fetched now, used by thread 0 at the very end
  if (r == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) cnt[i] = __ldcg(reinterpret_cast<const long long*>(a.buf.counters) + i);
  }
  int my_resets = 0, loco_rel = -1;
  const float4* part4 = reinterpret_cast<const float4*>(a.partials);
  for (int t0 = lo; t0 < hi; t0 += 4) {
    float4 v[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u) {                      // 12 independent 128-bit loads in flight
      const int t = min(t0 + u, hi - 1);
#pragma unroll
      for (int k = 0; k < 3; ++k) v[u][k] = __ldcg(part4 + (size_t)t * (STATS / 4) + k);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u;
      if (t < hi) {
        const float f[12] = {v[u][0].x, v[u][0].y, v[u][0].z, v[u][0].w, v[u][1].x, v[u][1].y, v[u][1].z, v[u][1].w,
                             v[u][2].x, v[u][2].y, v[u][2].z, v[u][2].w};
#pragma unroll
        for (int i = 0; i < 11; ++i) acc[i] += (double)f[i];
        if (P.kind == LM_KIND_COTRAIN && t >= tph) { acc[11] += (double)f[9]; acc[12] += (double)f[10]; }
        if (P.kind == LM_KIND_COTRAIN && t == tph) loco_rel = my_resets;      // resets in the locomotion half
        a.tile_offsets[t] = my_resets;
        my_resets += (int)f[10];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    const double v = warp_sum_d(acc[i]);
    if (lane == 0) fs.red_d[warp][i] = v;
  }
  int incl = my_resets;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += o;
  }
  if (lane == 31) fs.warp_scan[warp] = incl;
  __syncthreads();
  int base = 0, total_resets = 0;
  for (int w = 0; w < kWarps; ++w) {
    const int c = fs.warp_scan[w];
    if (w < warp) base += c;
    total_resets += c;
  }
  base += incl - my_resets;
  if (lo < hi) a.scan_base[r] = base;
  if (loco_rel >= 0) a.buf.num_reset[1] = base + loco_rel;
  if (r == 0) {
    double tot[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) {
      tot[i] = 0.0;
      for (int w = 0; w < kWarps; ++w) tot[i] += fs.red_d[w][i];
    }
    a.buf.num_reset[0] = total_resets;
    if (P.kind == LM_KIND_LOCO) a.buf.num_reset[1] = total_resets;
    if (P.kind == LM_KIND_MANI) a.buf.num_reset[1] = 0;
    float* ex = a.buf.extras;
    const double inv_n = 1.0 / (double)N;
    for (int i = 0; i < 7; ++i) ex[i] = (float)(tot[i] * inv_n);
    ex[LM_EXTRA_NAN_COUNT] = (float)tot[7];
    ex[LM_EXTRA_BELOW_M50_COUNT] = (float)tot[8];
    int64_t* c = a.buf.counters;
    const int groups = (P.kind == LM_KIND_COTRAIN) ? 3 : 1;
    // group 0 = all envs, 1 = locomotion half, 2 = manipulation half (QPC:619-633 / JLV:758-802)
    const double goal[3] = {tot[9], tot[9] - tot[11], tot[11]};
    const double rst[3] = {tot[10], tot[10] - tot[12], tot[12]};
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      if (g >= groups) break;
      int64_t ns = cnt[2 * g], nr = cnt[2 * g + 1];
      if (nr > (int64_t)P.max_reset_counts) {
        ex[LM_EXTRA_SUCCESS_RATE + g] = (float)ns / (float)nr;
        ns = 0; nr = 0;
      }
      c[2 * g] = ns + (int64_t)llrint(goal[g]);
      c[2 * g + 1] = nr + (int64_t)llrint(rst[g]);
    }
    *a.ticket = 0;
    if (kSeparate) a.ticket[1] = 0u;
  }
}

// per-CTA partial sums -> partials[tile] (QPC:560,610,619-633 / JLV:758-802).  finish_publish reduces the CTA's stats and
// lets ONE thread store them, fence and take the arrival ticket; the caller keeps the other threads busy (flag stores,
// observation copy-out) while that round trip is in flight and then calls finish_join: with inline_finish the last
// CTA to arrive runs the finish pass.  Thread->tile mapping and tree shape are fixed, so `extras` are deterministic.
template <int kThreads>
__device__ __forceinline__ void finish_publish(const PostArgs& a, const float stats[STATS], FinishSmem& fs) {
  const int r = threadIdx.x, lane = r & 31, warp = r >> 5;
  constexpr int kWarps = kThreads / 32;
#pragma unroll
  for (int i = 0; i < 11; ++i) {
    float v = stats[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) fs.red_f[warp][i] = v;
  }
  __syncthreads();
  if (r == 0) {
    float4* dst = reinterpret_cast<float4*>(a.partials + (size_t)blockIdx.x * STATS);
    float v[STATS];
#pragma unroll
    for (int i = 0; i < STATS; ++i) {
      v[i] = 0.0f;
      if (i < 11) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) v[i] += fs.red_f[w][i];
      }
    }
#pragma unroll
    for (int k = 0; k < STATS / 4; ++k) __stcg(dst + k, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
    if (a.inline_finish) {
      __threadfence();
      fs.is_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    } else {
      fs.is_last = 0;
      if (blockIdx.x == 0) a.ticket[1] = 1u;              // partials are fresh: a separately launched finish pass consumes them ONCE
    }
    a.tile_counts[blockIdx.x] = (int)v[10];               // dense per-tile reset counts (read by k_pre after a deferred finish)
  }
}
template <int kThreads>
__device__ __forceinline__ void finish_join(const KernelParams& kp, const PostArgs& a, FinishSmem& fs) {
  __syncthreads();
  if (!fs.is_last) return;
  // No second fence: every other CTA fenced its partial sums (L2, st.cg) before its ticket increment, this CTA observed
  // the final ticket value before the barrier above, and the finish pass reads them with L2 loads (ld.cg).  A full fence
  // here would only wait for this CTA's own observation stores.
  finish_tail<kThreads, false>(kp, a, fs);
}
template <int kThreads>
__device__ __forceinline__ void finish_stats(const KernelParams& kp, const PostArgs& a, const float stats[STATS], FinishSmem& fs) {
  finish_publish<kThreads>(a, stats, fs);
  finish_join<kThreads>(kp, a, fs);
}

// stand-alone finish pass for grids too large for a 128-thread tail
__global__ void __launch_bounds__(FINISH_THREADS) k_post_finish(const KernelParams kp, const PostArgs a) {
  __shared__ FinishSmem fs;
  finish_tail<FINISH_THREADS, true>(kp, a, fs);
}

// K_pre: (1) reset mask -> rank -> compacted ascending id list (the `nonzero()` of QPC:206); (2) clamp + stash
// actions, take_action.  The reset bodies themselves run in k_reset_apply over the compacted list, so the ~3 % of
// envs that reset do not drag every warp through the (double-precision) goal resampling path.
// one tile of pre_physics_step (shared by k_pre: one CTA per tile, and k_pre_early: resident CTAs walking the tiles)
__device__ __forceinline__ void pre_tile(const KernelParams& kp, const PreArgs& a, const int tile_idx, int* warp_base, int* base_part, int* list) {
  const LmTaskParams& P = kp.p;
  const Tile tl = tile_of(kp, tile_idx);
  const int r = threadIdx.x;
  const int N = P.num_envs;
  float* ring_cur = a.buf.actions_ring + (size_t)a.parity * N * LM_NUM_ACT;
  const LmRobotViews& V = tl.is_mani ? a.mani : a.loco;

  const int e = tl.e0 + r, hr = tl.hr0 + r;
  // ---- every independent load first: the reset flag, this tile's rank base and the thread's three action chunks are all
  //      in flight before anything waits (the reset bodies below are a long dependent chain on a few lanes)
  long long flag64 = 0;
  if (r < tl.n) flag64 = a.buf.reset_buf[e];
  int tile_base = 0;
  if (a.tile_counts == nullptr) {
    tile_base = a.scan_base[tile_idx / kp.d.scan_per] + a.tile_offsets[tile_idx];
  } else {
    // rank base from the dense per-tile reset counts of the last post step: the tiles before this one, four per 128-bit load
    const int4* c4 = reinterpret_cast<const int4*>(a.tile_counts);
    for (int i = r; 4 * i < tile_idx; i += T) {
      const int4 c = __ldcg(c4 + i);
      const int t0 = 4 * i, lim = tile_idx;
      tile_base += c.x + (t0 + 1 < lim ? c.y : 0) + (t0 + 2 < lim ? c.z : 0) + (t0 + 3 < lim ? c.w : 0);
    }
  }
  const int nflat4 = tl.n * (LM_NUM_ACT / 4);
  const float* act_g = a.actions + (size_t)tl.e0 * LM_NUM_ACT;
  const bool have_actions = a.actions != nullptr;       // NULL: delivered by the producing kernel through an LmActionSink
  float4 av[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int i = r + k * T;
    av[k] = (have_actions && i < nflat4) ? ld_nc_f4(act_g + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }

  // ---- actions: clamp (vec_env_rlgames.py:60), stash as current_actions, take_action (robot.py:444-461)
  if (have_actions) {
    float* cur_g = ring_cur + (size_t)tl.e0 * LM_NUM_ACT;
    const float range = sub(P.action_limit, -P.action_limit);
    const float offset = mul(add(-P.action_limit, P.action_limit), 0.5f);
    const bool write_targets = (P.control_mode == LM_CONTROL_VELOCITY || P.control_mode == LM_CONTROL_POSITION) && V.targets;
    float* tgt_g = write_targets ? V.targets + (size_t)tl.hr0 * LM_NUM_ACT : nullptr;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = r + k * T;
      if (i >= nflat4) continue;
      float4 v = av[k];
      if (P.clamp_actions) {
        v.x = fminf(fmaxf(v.x, -P.clip_actions), P.clip_actions); v.y = fminf(fmaxf(v.y, -P.clip_actions), P.clip_actions);
        v.z = fminf(fmaxf(v.z, -P.clip_actions), P.clip_actions); v.w = fminf(fmaxf(v.w, -P.clip_actions), P.clip_actions);
      }
      reinterpret_cast<float4*>(cur_g)[i] = v;
      if (write_targets) {
        if (P.zero_actions) v = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 t;
        t.x = add(mul(mul(v.x, range), 0.5f), offset); t.y = add(mul(mul(v.y, range), 0.5f), offset);
        t.z = add(mul(mul(v.z, range), 0.5f), offset); t.w = add(mul(mul(v.w, range), 0.5f), offset);
        reinterpret_cast<float4*>(tgt_g)[i] = t;
      }
    }
    if (P.control_mode == LM_CONTROL_EFFORT && V.targets && r < tl.n) {                // robot.py:455-459
      float* eff = V.targets + (size_t)hr * LM_NUM_DOF;
      const float* avp = a.actions + (size_t)e * LM_NUM_ACT;
#pragma unroll
      for (int j = 0; j < LM_NUM_DOF; ++j) eff[j] = 0.0f;
#pragma unroll
      for (int j = 0; j < LM_NUM_ACT; ++j) {
        float v = avp[j];
        if (P.clamp_actions) v = fminf(fmaxf(v, -P.clip_actions), P.clip_actions);
        if (P.zero_actions) v = 0.0f;
        eff[P.dof_index[j]] = add(mul(mul(v, range), 0.5f), offset);
      }
    }
  }

  // ---- reset mask -> rank -> compacted ascending id list (the `nonzero()` of QPC:206); the reset bodies run on the first
  //      lanes only, so the ~3 % of envs that reset do not drag every warp through the (double-precision) goal resampling
  const int flag = flag64 != 0;
  unsigned b = __ballot_sync(0xffffffffu, flag);
  if (a.tile_counts != nullptr) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tile_base += __shfl_xor_sync(0xffffffffu, tile_base, off);
    if ((r & 31) == 0) base_part[r >> 5] = tile_base;
  }
  if ((r & 31) == 0) warp_base[r >> 5] = __popc(b);
  __syncthreads();
  if (a.tile_counts != nullptr) {
    tile_base = 0;
#pragma unroll
    for (int w = 0; w < T / 32; ++w) tile_base += base_part[w];
  }
  int local_rank = __popc(b & ((1u << (r & 31)) - 1u)), count = 0;
#pragma unroll
  for (int w = 0; w < T / 32; ++w) {
    if (w < (r >> 5)) local_rank += warp_base[w];
    count += warp_base[w];
  }
  if (a.tile_counts != nullptr && r == 0) {             // num_reset, which the inline finish pass writes otherwise
    if (tile_idx == kp.d.num_tiles - 1) {
      a.buf.num_reset[0] = tile_base + count;
      if (P.kind == LM_KIND_LOCO) a.buf.num_reset[1] = tile_base + count;
      if (P.kind == LM_KIND_MANI) a.buf.num_reset[1] = 0;
    }
    if (P.kind == LM_KIND_COTRAIN && tile_idx == kp.d.tiles_per_half) a.buf.num_reset[1] = tile_base;   // resets in the locomotion half
  }
  if (count == 0) return;
  if (flag) list[local_rank] = r;
  __syncthreads();
  {
    // two lanes per reset env: role 0 = goal + indicator pose (the double-precision chain), role 1 = defaults + history.
    // Lanes are dealt round-robin over the four warps so that the chains of one tile run side by side.
    float* ring_last = a.buf.actions_ring + (size_t)(a.parity ^ 1) * N * LM_NUM_ACT;
    const int slot = ((r & 31) << 2) | (r >> 5);      // lane-major order: slot 0..3 sit in four different warps
    for (int idx = slot; idx < 2 * count; idx += T) {
      const int lr = list[idx >> 1];
      reset_one_env(kp, a, V, ring_last, tl.is_mani, tl.e0 + lr, tl.hr0 + lr, tile_base + (idx >> 1), idx & 1);
    }
  }
}

__global__ void __launch_bounds__(T) k_pre(const KernelParams kp, const PreArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ int warp_base[T / 32];
  __shared__ int base_part[T / 32];
  __shared__ int list[T];
  if (blockIdx.x == (unsigned)kp.d.num_tiles) {         // the extra CTA of a deferred finish: extras / counters of the previous post step
    __shared__ FinishSmem fs;
    finish_tail<T, true>(kp, a, fs);
    return;
  }
  pre_tile(kp, a, (int)blockIdx.x, warp_base, base_part, list);
}

// The reset half of pre_physics_step (actions == NULL: they arrive through an LmActionSink) for lm_task_set_early_reset: nothing in
// it depends on the kernel launched before it — the policy forward, which only READS the post step's outputs and delivers the
// actions — so under programmatic dependent launch the whole half runs NEXT TO that kernel instead of behind it.  The grid has at
// most one CTA per SM (128 threads: fits beside the policy kernel's 544-thread CTA when that kernel is built for <= 80 registers —
// registers are allocated in units of four warps), so every CTA is resident at once; each CTA takes up to four tiles per pass: all
// their flags and rank bases are fetched together and all their reset bodies run side by side on the CTA's lanes, one dependent
// chain instead of one per tile.  Only the very last instruction waits for the predecessor, so that this grid's completion — which
// releases the next post step — still implies the predecessor's.  Precondition (documented at lm_task_set_early_reset): the
// predecessor triggers its dependents only after its own dependency wait, i.e. after the post step has completed (k_mlp_forward
// does).
constexpr int EARLY_TILES = 4;
__global__ void __launch_bounds__(T) k_pre_early(const KernelParams kp, const PreArgs a, const int tile_ctas) {
  pdl_launch_dependents();
  const LmTaskParams& P = kp.p;
  __shared__ int warp_cnt[EARLY_TILES][T / 32];
  __shared__ int base_part[EARLY_TILES][T / 32];
  __shared__ int list[EARLY_TILES][T];
  if ((int)blockIdx.x == tile_ctas) {                   // the extra CTA of a deferred finish
    __shared__ FinishSmem fs;
    finish_tail<T, true>(kp, a, fs);
  } else {
    const int r = threadIdx.x, lane = r & 31, warp = r >> 5, nt = kp.d.num_tiles;
    float* ring_last = a.buf.actions_ring + (size_t)(a.parity ^ 1) * P.num_envs * LM_NUM_ACT;
    for (int first = blockIdx.x; first < nt; first += EARLY_TILES * tile_ctas) {
      Tile tl[EARLY_TILES];
      long long flag64[EARLY_TILES];
      int base[EARLY_TILES];
      // ---- phase 1: every independent load of the pass first (flags, per-tile counts)
#pragma unroll
      for (int j = 0; j < EARLY_TILES; ++j) {
        const int tile = first + j * tile_ctas;
        flag64[j] = 0; base[j] = 0;
        if (tile >= nt) continue;
        tl[j] = tile_of(kp, tile);
        if (r < tl[j].n) flag64[j] = a.buf.reset_buf[tl[j].e0 + r];
        if (a.tile_counts == nullptr) {
          if (r == 0) base[j] = a.scan_base[tile / kp.d.scan_per] + a.tile_offsets[tile];
        } else {
          const int4* c4 = reinterpret_cast<const int4*>(a.tile_counts);
          for (int i = r; 4 * i < tile; i += T) {
            const int4 c = __ldcg(c4 + i);
            const int t0 = 4 * i;
            base[j] += c.x + (t0 + 1 < tile ? c.y : 0) + (t0 + 2 < tile ? c.z : 0) + (t0 + 3 < tile ? c.w : 0);
          }
        }
      }
      unsigned bal[EARLY_TILES];
#pragma unroll
      for (int j = 0; j < EARLY_TILES; ++j) {
        bal[j] = __ballot_sync(0xffffffffu, flag64[j] != 0);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) base[j] += __shfl_xor_sync(0xffffffffu, base[j], off);
        if (lane == 0) { warp_cnt[j][warp] = __popc(bal[j]); base_part[j][warp] = base[j]; }
      }
      __syncthreads();
      int count[EARLY_TILES], total = 0;
#pragma unroll
      for (int j = 0; j < EARLY_TILES; ++j) {
        int local_rank = __popc(bal[j] & ((1u << lane) - 1u));
        count[j] = 0; base[j] = 0;
#pragma unroll
        for (int w = 0; w < T / 32; ++w) {
          if (w < warp) local_rank += warp_cnt[j][w];
          count[j] += warp_cnt[j][w];
          base[j] += base_part[j][w];
        }
        total += count[j];
        if (flag64[j] != 0) list[j][local_rank] = r;
        const int tile = first + j * tile_ctas;
        if (a.tile_counts != nullptr && r == 0 && tile < nt) {      // num_reset, which the inline finish pass writes otherwise
          if (tile == nt - 1) {
            a.buf.num_reset[0] = base[j] + count[j];
            if (P.kind == LM_KIND_LOCO) a.buf.num_reset[1] = base[j] + count[j];
            if (P.kind == LM_KIND_MANI) a.buf.num_reset[1] = 0;
          }
          if (P.kind == LM_KIND_COTRAIN && tile == kp.d.tiles_per_half) a.buf.num_reset[1] = base[j];
        }
      }
      __syncthreads();
      // ---- phase 2: the reset bodies of all tiles of the pass, two lanes per reset env (role 0 = goal + indicator pose, role 1 =
      //      defaults + history), dealt round-robin over the four warps
      const int slot = (lane << 2) | warp;
      for (int idx = slot; idx < 2 * total; idx += T) {
        int k = idx, j = 0;
#pragma unroll
        for (int jj = 0; jj < EARLY_TILES - 1; ++jj)
          if (j == jj && k >= 2 * count[jj]) { k -= 2 * count[jj]; j = jj + 1; }
        Tile tj = tl[0]; int bj = base[0];
#pragma unroll
        for (int jj = 1; jj < EARLY_TILES; ++jj)
          if (j == jj) { tj = tl[jj]; bj = base[jj]; }
        const int lr = list[j][k >> 1];
        reset_one_env(kp, a, tj.is_mani ? a.mani : a.loco, ring_last, tj.is_mani, tj.e0 + lr, tj.hr0 + lr, bj + (k >> 1), k & 1);
      }
      __syncthreads();                                  // the shared scratch is reused by the next pass
    }
  }
  pdl_wait();
}


__device__ __forceinline__ void ld12(const float* p, float v[12]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 x = reinterpret_cast<const float4*>(p)[k];
    v[4 * k] = x.x; v[4 * k + 1] = x.y; v[4 * k + 2] = x.z; v[4 * k + 3] = x.w;
  }
}
__device__ __forceinline__ void st12(float* p, const float v[12]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) reinterpret_cast<float4*>(p)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}
__device__ __forceinline__ Vec3 ld3(const float* p) { return Vec3{p[0], p[1], p[2]}; }

// ------------------------------------------------------------------------------------------------
// K_post, lean tiling (states alias obs: every co-train task).  One env per thread, 128 envs per CTA.  Only the
// arrays whose row stride defeats coalescing (20-wide DOF rows, 12-wide tip rows, 24-wide knee rows) are staged through
// shared memory by TMA; 16-byte-friendly rows are loaded straight into registers while the bulk copies fly.
// Observation rows leave through shared memory too: a thread storing its own 256-byte row would touch 32 different
// sectors per 128-bit store instruction (measured: ~45 % of the kernel's compute phase).  Each warp therefore parks
// its 32 rows, 16-byte chunks XOR-swizzled, in the staging space of its OWN 32 envs' inputs (dead once the limb half
// has consumed them — no CTA-wide barrier needed) and copies them out two full rows per instruction.
// 39 KB of shared memory per CTA -> 4 resident CTAs (16 warps) per SM; 65,536 envs are a single wave.
// ------------------------------------------------------------------------------------------------
struct LeanSmem {
  static constexpr int kDofPos = 0;
  static constexpr int kDofVel = kDofPos + T * 20;
  static constexpr int kTips = kDofVel + T * 20;
  static constexpr int kKnees = kTips + T * 12;
  static constexpr int kEnd = kKnees + T * 24;
  static constexpr int kBytes = kEnd * 4 + 64;
};
static_assert(LeanSmem::kDofVel % 32 == 0 && LeanSmem::kTips % 32 == 0 && LeanSmem::kKnees % 32 == 0, "staging rows must start on bank 0");

// float offset (in the CTA's dynamic shared memory) of observation staging row i (0..31) of warp w: the warp's own
// input rows hold 10 + 10 + 6 + 12 rows of 64 floats
__device__ __forceinline__ int obs_stage_row(int w, int i) {
  if (i < 10) return LeanSmem::kDofPos + w * 640 + i * 64;
  if (i < 20) return LeanSmem::kDofVel + w * 640 + (i - 10) * 64;
  if (i < 26) return LeanSmem::kTips + w * 384 + (i - 20) * 64;
  return LeanSmem::kKnees + w * 768 + (i - 26) * 64;
}

// Emits 16-byte chunk k of this lane's observation row into its staging row at chunk slot (k ^ swz), swz = lane & 15:
// eight consecutive lanes hit eight different bank groups (conflict-free 128-bit stores), and so does the copy-out
// that reads the rows back two at a time.
struct ObsToStage {
  float* row; int swz; float clip; int clamp;
  __device__ __forceinline__ void operator()(int k, float a, float b, float c, float d) const {
    if (clamp) {
      a = fminf(fmaxf(a, -clip), clip); b = fminf(fmaxf(b, -clip), clip);
      c = fminf(fmaxf(c, -clip), clip); d = fminf(fmaxf(d, -clip), clip);
    }
    *reinterpret_cast<float4*>(row + ((k ^ swz) << 2)) = make_float4(a, b, c, d);
  }
};

template <bool kStamp>
__global__ void __launch_bounds__(T, 4) k_post_lean(const KernelParams kp, const PostArgs a) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bar;
  __shared__ FinishSmem fs;
  const LmTaskParams& P = kp.p;
  const Tile tl = tile_of(kp, blockIdx.x);
  const int r = threadIdx.x, lane = r & 31, warp = r >> 5;
  const int N = P.num_envs;
  const LmRobotViews& V = tl.is_mani ? a.mani : a.loco;
  const float* ring_cur = a.buf.actions_ring + (size_t)a.parity * N * LM_NUM_ACT;
  const float* ring_last = a.buf.actions_ring + (size_t)(a.parity ^ 1) * N * LM_NUM_ACT;
  const int n = tl.n, h0 = tl.hr0, e0 = tl.e0;

  pdl_launch_dependents();
  stamp<kStamp>(a, 0);
  if (r == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  pdl_wait();
  Stager st{&bar, 0u, (n % 4) == 0};
  st.load(sm + LeanSmem::kDofPos, V.dof_pos + (size_t)h0 * 20, n * 20);
  st.load(sm + LeanSmem::kDofVel, V.dof_vel + (size_t)h0 * 20, n * 20);
  st.load(sm + LeanSmem::kTips, V.tip_pos + (size_t)h0 * 12, n * 12);
  st.load(sm + LeanSmem::kKnees, V.knee_pos + (size_t)h0 * 24, n * 24);
  if (st.use_tma && r == 0) mbar_expect_tx(&bar, st.bytes);

  EnvOut out;
  const bool valid = r < n;
  const int rr = valid ? r : 0;                       // padding lanes of a ragged tile recompute env 0 (results discarded)
  const int e = e0 + rr, hr = h0 + rr;
  float jp[12], jv[12], last_jv[12], cur[12], last[12], tips[12];
  PoseIn pin;
  {
    // register loads overlap the bulk copies; the small pose rows go first so the pose half can start early
    const float4 g4 = *reinterpret_cast<const float4*>(a.buf.goal_quaternions + (size_t)e * 4);
    const float4 q4 = *reinterpret_cast<const float4*>((tl.is_mani ? V.plate_quat : V.root_quat) + (size_t)hr * 4);
    pin.goal = Quat{g4.x, g4.y, g4.z, g4.w};
    pin.q = Quat{q4.x, q4.y, q4.z, q4.w};
    pin.O = ld3(a.env_pos + (size_t)hr * 3);
    pin.Pw = ld3((tl.is_mani ? V.plate_pos : V.root_pos) + (size_t)hr * 3);
    pin.lin = ld3((tl.is_mani ? V.plate_lin_vel : V.root_lin_vel) + (size_t)hr * 3);
    pin.ang = ld3((tl.is_mani ? V.plate_ang_vel : V.root_ang_vel) + (size_t)hr * 3);
    pin.progress = (int)a.buf.progress_buf[e];
    pin.reset = (int)a.buf.reset_buf[e];
    pin.succ_prev = (int)a.buf.successes[e];
    pin.consec = (int)a.buf.consecutive_successes[e];
    ld12(ring_cur + (size_t)e * 12, cur);
    ld12(ring_last + (size_t)e * 12, last);
    ld12(a.buf.last_joint_vel + (size_t)e * 12, last_jv);
  }
  stamp<kStamp>(a, 1);
  if (st.use_tma) mbar_wait(&bar, 0); else __syncthreads();
  stamp<kStamp>(a, 2);
  if (kp.d.dof_identity) {
    ld12(sm + LeanSmem::kDofPos + rr * 20, jp);
    ld12(sm + LeanSmem::kDofVel + rr * 20, jv);
  } else {
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      jp[j] = sm[LeanSmem::kDofPos + rr * 20 + P.dof_index[j]];
      jv[j] = sm[LeanSmem::kDofVel + rr * 20 + P.dof_index[j]];
    }
  }
  ld12(sm + LeanSmem::kTips + rr * 12, tips);
  if (valid) st12(a.buf.last_joint_vel + (size_t)e * 12, jv);                          // robot.py:291
  if (valid && a.pd_joint_pos != nullptr) {
    st12(a.pd_joint_pos + (size_t)e * 12, jp);
    st12(a.pd_joint_vel + (size_t)e * 12, jv);
  }
  LimbOut lo;
  FrameShare f;
  const Vec3 pl = pose_frame(pin, f);
  // the frame part of the limb half first: it consumes the knees, the last staged input, before the observation chunks
  // start to overwrite this warp's staging space
  limb_frame(kp, tl.is_mani, pin.O, f, tips, sm + LeanSmem::kKnees + rr * 24, lo);
  if (valid) st12(a.buf.last_base_tip + (size_t)e * 12, tips);                         // QPC:426
  __syncwarp();                                        // every lane of the warp is done with the staged inputs
  PoseMid m;
  {
    const ObsToStage sink{sm + obs_stage_row(warp, lane), lane & 15, P.clip_observations, a.clamp_observations};
    limb_joints(kp, jp, jv, last_jv, cur, last, sink, lo, a.pd_targets ? a.pd_targets + (size_t)e * 12 : nullptr,
                a.pd_last_targets ? a.pd_last_targets + (size_t)e * 12 : nullptr, a.pd_target_obs_scale);
    pose_main(kp, tl.is_mani, pin, pl, f, sink, out.cur_q, m);
  }
  pose_finish(kp, pin, m, lo, out);
  if (!valid) {
#pragma unroll
    for (int i = 0; i < STATS; ++i) out.stats[i] = 0.0f;
  }
  // ---- publish the CTA's partial sums (one thread: store, fence, ticket) while everybody else stores
  finish_publish<T>(a, out.stats, fs);
  if (valid) store_flags(a.buf, e, out);
  __syncwarp();
  {
    // copy-out: instruction j moves rows 2j and 2j+1 of the warp (2 x 256 contiguous bytes)
    const int half = lane >> 4, c = lane & 15;
    float4* obs_g = reinterpret_cast<float4*>(a.buf.obs_buf + (size_t)(e0 + (warp << 5)) * LM_NUM_OBS);
    const int rows_here = min(32, n - (warp << 5));
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int row = 2 * j + half;
      const float4 v = *reinterpret_cast<const float4*>(sm + obs_stage_row(warp, row) + ((c ^ (row & 15)) << 2));
      if (row < rows_here) obs_g[row * 16 + c] = v;
    }
  }
  stamp<kStamp>(a, 3);
  finish_join<T>(kp, a, fs);
  stamp<kStamp>(a, 4);
}

// ------------------------------------------------------------------------------------------------
// K_post, full tiling (single-task files with the 93-wide states_buf, QPC:406-423): everything staged in
// shared memory so the 93-wide rows can be gathered and stored coalesced.
// ------------------------------------------------------------------------------------------------
struct PostSmem {
  // front: inputs that are dead once the env math has run; the [T][93] states rows are assembled over them afterwards
  static constexpr int kDofPos = 0;
  static constexpr int kDofVel = kDofPos + T * 20;
  static constexpr int kPos = kDofVel + T * 20;      // root_pos | plate_pos
  static constexpr int kQuat = kPos + T * 3;
  static constexpr int kLin = kQuat + T * 4;
  static constexpr int kAng = kLin + T * 3;
  static constexpr int kKnees = kAng + T * 3;
  static constexpr int kEnvPos = kKnees + T * 24;
  static constexpr int kCur = kEnvPos + T * 3;
  static constexpr int kLast = kCur + T * 12;
  static constexpr int kFrontEnd = kLast + T * 12;
  static constexpr int kStates = 0;                  // [T][93] rows, contiguous like the global tile span
  // back: what the write-back phase still reads
  static constexpr int kTips = kFrontEnd;
  static constexpr int kGoal = kTips + T * 12;
  static constexpr int kLastJv = kGoal + T * 4;
  static constexpr int kLastTip = kLastJv + T * 12;
  static constexpr int kCurQ = kLastTip + T * 12;
  static constexpr int kObs = kCurQ + T * 4;
  static constexpr int kEnd = kObs + T * OBS_PAD;
  static constexpr int kBytes = kEnd * 4 + 64;
};
static_assert(PostSmem::kFrontEnd >= T * LM_NUM_STATES_FULL, "the states rows fit the dead input staging");
static_assert(T % 4 == 0, "tile spans must stay 16-byte aligned");

// the 93-wide states row of env r (QPC:406-423), assembled from what the env math left in shared memory: 93 conflict-free
// scalar stores per thread (row stride 93 words is odd) instead of a per-element gather with a division by 93
__device__ __forceinline__ void states_row(const LmTaskParams& P, float* sm, int r, bool clamp) {
  const float* obs = sm + PostSmem::kObs + r * OBS_PAD;
  float* row = sm + PostSmem::kStates + r * LM_NUM_STATES_FULL;
  const float lim = P.clip_observations;
  auto put = [&](int c, float v) { row[c] = clamp ? fminf(fmaxf(v, -lim), lim) : v; };
#pragma unroll
  for (int k = 0; k < 3; ++k) put(k, obs[k]);                                                     // scaled position
#pragma unroll
  for (int k = 0; k < 3; ++k) put(3 + k, obs[10 + k]);                                            // scaled linear velocity
#pragma unroll
  for (int k = 0; k < 4; ++k) put(6 + k, mul(P.quaternion_obs_scale, sm[PostSmem::kCurQ + r * 4 + k]));
#pragma unroll
  for (int k = 0; k < 3; ++k) put(10 + k, obs[13 + k]);                                           // scaled angular velocity
#pragma unroll
  for (int k = 0; k < 24; ++k) put(13 + k, obs[16 + k]);                                          // joint positions, velocities
#pragma unroll
  for (int k = 0; k < 4; ++k) put(37 + k, mul(P.quaternion_obs_scale, sm[PostSmem::kGoal + r * 4 + k]));
#pragma unroll
  for (int k = 0; k < 4; ++k) put(41 + k, obs[6 + k]);                                            // quat diff
#pragma unroll
  for (int k = 0; k < 12; ++k) put(45 + k, mul(P.tip_position_scale, sm[PostSmem::kTips + r * 12 + k]));      // tips now (base frame)
#pragma unroll
  for (int k = 0; k < 12; ++k) put(57 + k, mul(P.tip_position_scale, sm[PostSmem::kLastTip + r * 12 + k]));   // tips previous step
#pragma unroll
  for (int k = 0; k < 24; ++k) put(69 + k, obs[40 + k]);                                          // current, last actions
}

__global__ void __launch_bounds__(T, 2) k_post(const KernelParams kp, const PostArgs a) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bar;
  __shared__ FinishSmem fs;
  const LmTaskParams& P = kp.p;
  const Tile tl = tile_of(kp, blockIdx.x);
  const int r = threadIdx.x;
  const int N = P.num_envs;
  const LmRobotViews& V = tl.is_mani ? a.mani : a.loco;
  const float* ring_cur = a.buf.actions_ring + (size_t)a.parity * N * LM_NUM_ACT;
  const float* ring_last = a.buf.actions_ring + (size_t)(a.parity ^ 1) * N * LM_NUM_ACT;
  const bool full_states = (P.num_states == LM_NUM_STATES_FULL) && a.buf.states_buf != nullptr;

  if (r == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  Stager st{&bar, 0u, (tl.n % 4) == 0};
  const int n = tl.n, h0 = tl.hr0, e0 = tl.e0;
  st.load(sm + PostSmem::kDofPos, V.dof_pos + (size_t)h0 * 20, n * 20);
  st.load(sm + PostSmem::kDofVel, V.dof_vel + (size_t)h0 * 20, n * 20);
  st.load(sm + PostSmem::kPos, (tl.is_mani ? V.plate_pos : V.root_pos) + (size_t)h0 * 3, n * 3);
  st.load(sm + PostSmem::kQuat, (tl.is_mani ? V.plate_quat : V.root_quat) + (size_t)h0 * 4, n * 4);
  st.load(sm + PostSmem::kLin, (tl.is_mani ? V.plate_lin_vel : V.root_lin_vel) + (size_t)h0 * 3, n * 3);
  st.load(sm + PostSmem::kAng, (tl.is_mani ? V.plate_ang_vel : V.root_ang_vel) + (size_t)h0 * 3, n * 3);
  st.load(sm + PostSmem::kTips, V.tip_pos + (size_t)h0 * 12, n * 12);
  st.load(sm + PostSmem::kKnees, V.knee_pos + (size_t)h0 * 24, n * 24);
  st.load(sm + PostSmem::kEnvPos, a.env_pos + (size_t)h0 * 3, n * 3);
  st.load(sm + PostSmem::kGoal, a.buf.goal_quaternions + (size_t)e0 * 4, n * 4);
  st.load(sm + PostSmem::kCur, ring_cur + (size_t)e0 * 12, n * 12);
  st.load(sm + PostSmem::kLast, ring_last + (size_t)e0 * 12, n * 12);
  st.load(sm + PostSmem::kLastJv, a.buf.last_joint_vel + (size_t)e0 * 12, n * 12);
  if (full_states) st.load(sm + PostSmem::kLastTip, a.buf.last_base_tip + (size_t)e0 * 12, n * 12);
  if (st.use_tma && r == 0) mbar_expect_tx(&bar, st.bytes);
  EnvIn in;
  EnvOut out;
  in.progress = in.reset = in.succ_prev = in.consec = 0;
  if (r < n) {   // int64 flags straight from global while the bulk copies fly
    in.progress = (int)a.buf.progress_buf[e0 + r];
    in.reset = (int)a.buf.reset_buf[e0 + r];
    in.succ_prev = (int)a.buf.successes[e0 + r];
    in.consec = (int)a.buf.consecutive_successes[e0 + r];
  }
  if (st.use_tma) mbar_wait(&bar, 0); else __syncthreads();

#pragma unroll
  for (int i = 0; i < STATS; ++i) out.stats[i] = 0.0f;
  if (r < n) {
    if (kp.d.dof_identity) {
      ld12(sm + PostSmem::kDofPos + r * 20, in.jp);
      ld12(sm + PostSmem::kDofVel + r * 20, in.jv);
    } else {
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        in.jp[j] = sm[PostSmem::kDofPos + r * 20 + P.dof_index[j]];
        in.jv[j] = sm[PostSmem::kDofVel + r * 20 + P.dof_index[j]];
      }
    }
    ld12(sm + PostSmem::kCur + r * 12, in.cur);
    ld12(sm + PostSmem::kLast + r * 12, in.last);
    ld12(sm + PostSmem::kLastJv + r * 12, in.last_jv);
    in.O = ld3(sm + PostSmem::kEnvPos + r * 3);
    in.Pw = ld3(sm + PostSmem::kPos + r * 3);
    in.lin = ld3(sm + PostSmem::kLin + r * 3);
    in.ang = ld3(sm + PostSmem::kAng + r * 3);
    const float4 qv = *reinterpret_cast<const float4*>(sm + PostSmem::kQuat + r * 4);
    const float4 gv = *reinterpret_cast<const float4*>(sm + PostSmem::kGoal + r * 4);
    in.q = Quat{qv.x, qv.y, qv.z, qv.w};
    in.goal = Quat{gv.x, gv.y, gv.z, gv.w};
    st12(sm + PostSmem::kLastJv + r * 12, in.jv);                                        // last_joint_vel <- jv (robot.py:291)
    env_math(kp, tl.is_mani, in, sm + PostSmem::kTips + r * 12, sm + PostSmem::kKnees + r * 24,
             ObsToGlobal{reinterpret_cast<float4*>(sm + PostSmem::kObs + r * OBS_PAD), P.clip_observations, a.clamp_observations}, out);
    float* cq = sm + PostSmem::kCurQ + r * 4;
    cq[0] = out.cur_q.w; cq[1] = out.cur_q.x; cq[2] = out.cur_q.y; cq[3] = out.cur_q.z;
    store_flags(a.buf, e0 + r, out);
  }
  fence_proxy_async();
  __syncthreads();
  // ---- write back: obs rows (coalesced 128-bit), states gather, last_joint_vel / last_base_tip (TMA bulk store)
  {
    float* obs_g = a.buf.obs_buf + (size_t)e0 * LM_NUM_OBS;
    for (int i = r; i < n * 16; i += T) {
      const int row = i >> 4, c4 = i & 15;
      reinterpret_cast<float4*>(obs_g)[i] = *reinterpret_cast<const float4*>(sm + PostSmem::kObs + row * OBS_PAD + c4 * 4);
    }
    if (full_states) {
      // rows assembled over the dead input staging (every thread passed the barrier above: nobody reads it any more), then the
      // tile's contiguous [n][93] span leaves with 128-bit stores
      if (r < n) states_row(P, sm, r, a.clamp_observations != 0);
      __syncthreads();
      float* st_g = a.buf.states_buf + (size_t)e0 * LM_NUM_STATES_FULL;
      const int total_el = n * LM_NUM_STATES_FULL;
      const float* rows = sm + PostSmem::kStates;
      if ((total_el & 3) == 0 && ((e0 * LM_NUM_STATES_FULL) & 3) == 0) {
        for (int i = r; i < (total_el >> 2); i += T) reinterpret_cast<float4*>(st_g)[i] = reinterpret_cast<const float4*>(rows)[i];
      } else {
        for (int i = r; i < total_el; i += T) st_g[i] = rows[i];
      }
    }
    if (st.use_tma) {
      if (r == 0) {
        tma_store_1d(a.buf.last_joint_vel + (size_t)e0 * 12, sm + PostSmem::kLastJv, (uint32_t)n * 48u);
        tma_store_1d(a.buf.last_base_tip + (size_t)e0 * 12, sm + PostSmem::kTips, (uint32_t)n * 48u);
        tma_store_commit();
      }
    } else {
      for (int i = r; i < n * 12; i += T) {
        a.buf.last_joint_vel[(size_t)e0 * 12 + i] = sm[PostSmem::kLastJv + i];
        a.buf.last_base_tip[(size_t)e0 * 12 + i] = sm[PostSmem::kTips + i];
      }
    }
  }
  finish_stats<T>(kp, a, out.stats, fs);
  if (st.use_tma && r == 0) tma_store_wait_all();
}

// clamp + copy (_process_data, envs/vec_env_rlgames.py:41-46)
__global__ void k_clamp_copy(const float* __restrict__ in, float* __restrict__ out, int64_t count, float clip) {
  const int64_t n4 = count >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(in)[i];
    v.x = fminf(fmaxf(v.x, -clip), clip); v.y = fminf(fmaxf(v.y, -clip), clip);
    v.z = fminf(fmaxf(v.z, -clip), clip); v.w = fminf(fmaxf(v.w, -clip), clip);
    reinterpret_cast<float4*>(out)[i] = v;
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = fminf(fmaxf(in[i], -clip), clip);
}

}  // namespace lm

// =================================================================================================
// host side
// =================================================================================================
using namespace lm;

struct LmTask {
  KernelParams kp;
  int device;
  int parity;
  int sm_count;
  int32_t* tile_counts;
  int32_t* tile_offsets;
  int32_t* scan_base;
  int inline_finish;
  int defer_finish;                // the caller launches the finish pass itself (lm_task_post_finish), e.g. on a second stream
  int finish_pending;              // a deferred finish pass has not run yet
  int early_reset;                 // lm_task_set_early_reset: the reset half of pre_step may run beside the kernel launched before it
  LmPdState pd;
  int has_pd;
  float* partials;
  unsigned int* tickets;     // [4]: scan ticket, post ticket, 'partials are fresh' flag (deferred finish), spare
  int clamp_observations;
  long long* timeline;
  int scan_valid;            // tile_offsets / num_reset were produced by the last post_step and reset_buf is untouched since
  // host-buffer step with a dedicated copy stream (LmHostStep.h2d_stream): state landed / state arena free for the next copy
  cudaEvent_t ev_state, ev_consumed;
  int have_events;
};

static void host_rot_from_quat(const float q[4], float m[9]) {   // same formulas as lm::rot_from_quat_scipy
  double w = q[0], x = q[1], y = q[2], z = q[3];
  double n = sqrt(x * x + y * y + z * z + w * w);
  x /= n; y /= n; z /= n; w /= n;
  double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w, xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
  m[0] = (float)(x2 - y2 - z2 + w2); m[1] = (float)(2.0 * (xy - zw)); m[2] = (float)(2.0 * (xz + yw));
  m[3] = (float)(2.0 * (xy + zw)); m[4] = (float)(-x2 + y2 - z2 + w2); m[5] = (float)(2.0 * (yz - xw));
  m[6] = (float)(2.0 * (xz - yw)); m[7] = (float)(2.0 * (yz + xw)); m[8] = (float)(-x2 - y2 + z2 + w2);
}
// volatile keeps the host compiler from contracting a*b+c into an FMA: the fp32 values must equal torch.bmm's
static float seq_dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
  volatile float p0 = a0 * b0, p1 = a1 * b1, p2 = a2 * b2;
  volatile float s = p0 + p1;
  volatile float s2 = s + p2;
  return s2;
}

extern "C" {

int lm_abi_version(void) { return LM_ABI_VERSION; }
const char* lm_last_error(void) { return lm_error_buffer(); }

int lm_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(LmTaskParams);
    case 1: return (int)sizeof(LmRobotViews);
    case 2: return (int)sizeof(LmTaskBuffers);
    case 3: return (int)sizeof(LmHostStep);
    case 4: return (int)sizeof(LmPdState);
    case 5: return (int)sizeof(LmRunningStats);
    case 6: return (int)sizeof(LmRolloutStep);
    case 7: return (int)sizeof(LmRolloutBuffers);
    case 8: return (int)sizeof(LmPpoBatch);
    case 9: return (int)sizeof(LmNoiseSpec);
    case 10: return (int)sizeof(LmNoiseState);
    case 11: return (int)sizeof(LmCustomController);
    case 12: return (int)sizeof(LmAdamSchedule);
    case 13: return (int)sizeof(LmActionSink);
    case 14: return (int)sizeof(LmHostPolicy);
    default: return -1;
  }
}

int lm_device_sm_count(int device) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
  return prop.multiProcessorCount;
}

LmTask* lm_task_create(const LmTaskParams* params, int device) {
  if (!params) { lm_set_error("lm_task_create: params is NULL"); return nullptr; }
  const LmTaskParams& p = *params;
  if (p.num_envs <= 0 || p.kind < 0 || p.kind > 2) { lm_set_error("lm_task_create: bad num_envs/kind"); return nullptr; }
  if (p.kind == LM_KIND_COTRAIN && (p.num_envs % 2)) { lm_set_error("Number of envs must be a multiplier of 2."); return nullptr; }
  if (p.num_states != LM_NUM_OBS && p.num_states != LM_NUM_STATES_FULL) { lm_set_error("num_states must be 64 or 93"); return nullptr; }
  cudaDeviceProp prop;
  cudaError_t err = cudaGetDeviceProperties(&prop, device);
  if (err != cudaSuccess) { lm_set_error("lm_task_create: no CUDA device (%s)", cudaGetErrorString(err)); return nullptr; }
  if (prop.major != 10) { lm_set_error("lm_task_create: device %d is sm_%d%d; this library is sm_100a only", device, prop.major, prop.minor); return nullptr; }
  if ((err = cudaSetDevice(device)) != cudaSuccess) { lm_set_error("cudaSetDevice: %s", cudaGetErrorString(err)); return nullptr; }
  LmTask* t = new LmTask();
  memset(t, 0, sizeof(*t));
  t->kp.p = p;
  t->device = device;
  t->sm_count = prop.multiProcessorCount;
  Derived& d = t->kp.d;
  d.half = (p.kind == LM_KIND_COTRAIN) ? p.num_envs / 2 : p.num_envs;
  d.tiles_per_half = (d.half + T - 1) / T;
  d.num_tiles = d.tiles_per_half * ((p.kind == LM_KIND_COTRAIN) ? 2 : 1);
  t->defer_finish = 0;
  t->inline_finish = d.num_tiles <= INLINE_FINISH_MAX_TILES;
  {
    const int threads = t->inline_finish ? T : FINISH_THREADS;
    d.scan_per = (d.num_tiles + threads - 1) / threads;
  }
  d.dof_identity = 1;
  for (int j = 0; j < LM_NUM_ACT; ++j) {
    if (p.dof_index[j] < 0 || p.dof_index[j] >= LM_NUM_DOF) { lm_set_error("dof_index[%d] out of range", j); delete t; return nullptr; }
    d.dof_identity &= (p.dof_index[j] == j);
  }
  host_rot_from_quat(p.mani_default_quaternion, d.R0);
  const float* p0 = p.mani_default_position;
  for (int i = 0; i < 3; ++i)   // bmm(R0^T, -p0)
    d.p0inv[i] = seq_dot3(d.R0[0 + i], -p0[0], d.R0[3 + i], -p0[1], d.R0[6 + i], -p0[2]);
  for (int k = 0; k < 4; ++k)
    for (int i = 0; i < 3; ++i) {
      volatile float s = seq_dot3(d.R0[3 * i], p.corners[k][0], d.R0[3 * i + 1], p.corners[k][1], d.R0[3 * i + 2], p.corners[k][2]);
      volatile float v = s + p0[i];
      d.corners_world[k][i] = v;
    }
  bool ok = cudaMalloc(&t->tile_counts, sizeof(int32_t) * d.num_tiles) == cudaSuccess &&
            cudaMalloc(&t->tile_offsets, sizeof(int32_t) * d.num_tiles) == cudaSuccess &&
            cudaMalloc(&t->scan_base, sizeof(int32_t) * FINISH_THREADS) == cudaSuccess &&
            cudaMalloc(&t->partials, sizeof(float) * STATS * d.num_tiles) == cudaSuccess &&
            cudaMalloc(&t->tickets, sizeof(unsigned int) * 4) == cudaSuccess;
  if (!ok) { lm_set_error("lm_task_create: cudaMalloc failed"); lm_task_destroy(t); return nullptr; }
  cudaMemset(t->tickets, 0, sizeof(unsigned int) * 4);
  cudaMemset(t->scan_base, 0, sizeof(int32_t) * FINISH_THREADS);
  err = cudaFuncSetAttribute(k_post_lean<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LeanSmem::kBytes);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(k_post_lean<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LeanSmem::kBytes);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(k_post, cudaFuncAttributeMaxDynamicSharedMemorySize, PostSmem::kBytes);
  if (err != cudaSuccess) { lm_set_error("cudaFuncSetAttribute(k_post): %s", cudaGetErrorString(err)); lm_task_destroy(t); return nullptr; }
  return t;
}

void lm_task_destroy(LmTask* t) {
  if (!t) return;
  cudaFree(t->tile_counts); cudaFree(t->tile_offsets); cudaFree(t->scan_base); cudaFree(t->partials); cudaFree(t->tickets);
  if (t->have_events) { cudaEventDestroy(t->ev_state); cudaEventDestroy(t->ev_consumed); }
  delete t;
}

int lm_task_step_parity(const LmTask* t) { return t ? t->parity : -1; }

int lm_task_set_clamp_observations(LmTask* t, int on) { if (!t) return LM_ERR_BAD_ARG; t->clamp_observations = on; return LM_OK; }

int lm_task_set_timeline(LmTask* t, long long* stamps_dev) { if (!t) return LM_ERR_BAD_ARG; t->timeline = stamps_dev; return LM_OK; }

int lm_task_set_pd_state(LmTask* t, const LmPdState* pd) {
  if (!t) return lm_fail(LM_ERR_BAD_ARG, "lm_task_set_pd_state: task is NULL");
  if (!pd) { memset(&t->pd, 0, sizeof(t->pd)); t->has_pd = 0; return LM_OK; }
  // With the 93-wide states variant (the *_custom_controller files) k_post keeps the action rows in obs[40:64] and does not
  // carry the post-step joint state: lm_cc_post adds the target rows and the carry.  k_pre resets the PD rows either way.
  const void* ptrs[5] = {pd->joint_pos, pd->joint_vel, pd->targets, pd->last_targets, pd->targets_se};
  for (int i = 0; i < 5; ++i) {
    if (!ptrs[i]) return lm_fail(LM_ERR_BAD_ARG, "lm_task_set_pd_state: NULL tensor");
    if (((uintptr_t)ptrs[i]) & 15) return lm_fail(LM_ERR_ALIGNMENT, "lm_task_set_pd_state: tensor not 16-byte aligned");
  }
  t->pd = *pd; t->has_pd = 1;
  return LM_OK;
}

int lm_task_invalidate_reset_scan(LmTask* t) { if (!t) return LM_ERR_BAD_ARG; t->scan_valid = 0; return LM_OK; }

static int check_views(const LmTask* t, const LmRobotViews* loco, const LmRobotViews* mani, const float* env_pos,
                       const LmTaskBuffers* buf) {
  const int kind = t->kp.p.kind;
  if (!buf || !env_pos) return lm_fail(LM_ERR_BAD_ARG, "buffers/env_pos NULL");
  if ((kind != LM_KIND_MANI) && !loco) return lm_fail(LM_ERR_BAD_ARG, "locomotion views required");
  if ((kind != LM_KIND_LOCO) && !mani) return lm_fail(LM_ERR_BAD_ARG, "manipulation views required");
#define LM_CHK(ptr)                                                                                   \
  do {                                                                                                \
    if (!(ptr)) return lm_fail(LM_ERR_BAD_ARG, #ptr " is NULL");                                      \
    if (((uintptr_t)(ptr)) & 15) return lm_fail(LM_ERR_ALIGNMENT, #ptr " is not 16-byte aligned");    \
  } while (0)
  LM_CHK(env_pos);
  LM_CHK(buf->obs_buf); LM_CHK(buf->rew_buf); LM_CHK(buf->reset_buf); LM_CHK(buf->progress_buf); LM_CHK(buf->successes);
  LM_CHK(buf->consecutive_successes); LM_CHK(buf->goal_reset_buf); LM_CHK(buf->actions_ring); LM_CHK(buf->last_joint_vel);
  LM_CHK(buf->last_base_tip); LM_CHK(buf->goal_quaternions); LM_CHK(buf->reset_env_ids); LM_CHK(buf->num_reset);
  LM_CHK(buf->extras); LM_CHK(buf->counters);
  if (t->kp.p.num_states == LM_NUM_STATES_FULL) LM_CHK(buf->states_buf);
  if (loco && kind != LM_KIND_MANI) {
    LM_CHK(loco->dof_pos); LM_CHK(loco->dof_vel); LM_CHK(loco->root_pos); LM_CHK(loco->root_quat); LM_CHK(loco->root_lin_vel);
    LM_CHK(loco->root_ang_vel); LM_CHK(loco->tip_pos); LM_CHK(loco->knee_pos); LM_CHK(loco->indicator_pos); LM_CHK(loco->indicator_quat);
  }
  if (mani && kind != LM_KIND_LOCO) {
    LM_CHK(mani->dof_pos); LM_CHK(mani->dof_vel); LM_CHK(mani->tip_pos); LM_CHK(mani->knee_pos); LM_CHK(mani->plate_pos);
    LM_CHK(mani->plate_quat); LM_CHK(mani->plate_lin_vel); LM_CHK(mani->plate_ang_vel); LM_CHK(mani->indicator_pos); LM_CHK(mani->indicator_quat);
  }
#undef LM_CHK
  return LM_OK;
}

// CUDA-graph replays do not run the host side of lm_task_pre_step: a caller replaying captured steps tells the handle which ring
// slot the replayed step used, so that host-side readers (current_actions / last_actions, lm_task_action_sink) stay in step
int lm_task_set_step_parity(LmTask* t, int parity) {
  if (!t || (parity != 0 && parity != 1)) return lm_fail(LM_ERR_BAD_ARG, "lm_task_set_step_parity: bad argument");
  t->parity = parity;
  return LM_OK;
}

int lm_task_action_sink(LmTask* t, const LmRobotViews* loco, const LmRobotViews* mani, const LmTaskBuffers* buf, LmActionSink* out) {
  if (!t || !buf || !out) return lm_fail(LM_ERR_BAD_ARG, "lm_task_action_sink: NULL argument");
  const LmTaskParams& P = t->kp.p;
  if (!(P.control_mode == LM_CONTROL_VELOCITY || P.control_mode == LM_CONTROL_POSITION))
    return lm_fail(LM_ERR_UNSUPPORTED, "lm_task_action_sink: velocity / position control only");
  if (!buf->actions_ring) return lm_fail(LM_ERR_BAD_ARG, "lm_task_action_sink: NULL actions_ring");
  memset(out, 0, sizeof(*out));
  out->ring_next = buf->actions_ring + (size_t)(t->parity ^ 1) * P.num_envs * LM_NUM_ACT;    // the next pre_step flips the parity
  out->targets_loco = (loco && P.kind != LM_KIND_MANI) ? loco->targets : nullptr;
  out->targets_mani = (mani && P.kind != LM_KIND_LOCO) ? mani->targets : nullptr;
  out->loco_envs = P.kind == LM_KIND_MANI ? 0 : (P.kind == LM_KIND_COTRAIN ? t->kp.d.half : P.num_envs);
  out->clamp_actions = P.clamp_actions; out->zero_actions = P.zero_actions;
  out->clip_actions = P.clip_actions; out->action_limit = P.action_limit;
  return LM_OK;
}

int lm_task_pre_step(LmTask* t, const LmRobotViews* loco, const LmRobotViews* mani, const float* env_pos,
                     const LmTaskBuffers* buf, const float* actions, const float* uniforms, void* stream) {
  if (!t || !uniforms) return lm_fail(LM_ERR_BAD_ARG, "lm_task_pre_step: NULL argument");
  // actions == NULL: the producing kernel delivered them through an LmActionSink (lm_task_action_sink): reset half only
  if (!actions && (t->has_pd || t->kp.p.control_mode == LM_CONTROL_EFFORT))
    return lm_fail(LM_ERR_BAD_ARG, "lm_task_pre_step: actions == NULL needs a velocity / position control task (action sink)");
  if (((uintptr_t)actions) & 15) return lm_fail(LM_ERR_ALIGNMENT, "actions is not 16-byte aligned");
  int rc = check_views(t, loco, mani, env_pos, buf);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const Derived& d = t->kp.d;
  t->parity ^= 1;
  const bool scan_was_valid = t->scan_valid != 0;
  if (!scan_was_valid)  // first step, or reset_buf was written from outside (task.reset(), lm_task_invalidate_reset_scan)
    k_reset_count<<<d.num_tiles, T, 0, s>>>(t->kp, buf->reset_buf, t->tile_counts, t->tile_offsets, t->scan_base, buf->num_reset, t->tickets + 0);
  t->scan_valid = 0;
  PreArgs a;
  memset(&a, 0, sizeof(a));
  if (loco) a.loco = *loco;
  if (mani) a.mani = *mani;
  a.env_pos = env_pos; a.buf = *buf; a.actions = actions; a.uniforms = uniforms;
  a.tile_offsets = t->tile_offsets; a.scan_base = t->scan_base; a.parity = t->parity;
  if (t->has_pd) a.pd = t->pd;
  int grid = d.num_tiles;
  bool finish_behind = false;
  if (t->defer_finish) {
    // the reset scan: k_reset_count's (absolute tile_offsets, scan_per of the 128-thread layout) when it just ran, else from the
    // per-tile counts the last post step left.  The pending finish pass: an extra CTA of this launch (<= 1024 tiles: a 128-thread
    // pass is short), else the stand-alone 1024-thread kernel behind it.
    if (scan_was_valid) a.tile_counts = t->tile_counts;
    if (t->finish_pending) {
      // (with k_reset_count's scan in use the pass must not run beside the reset bodies: it rewrites tile_offsets / scan_base)
      if (d.num_tiles <= INLINE_FINISH_MAX_TILES && scan_was_valid) {
        a.partials = t->partials; a.ticket = t->tickets + 1;
        grid = d.num_tiles + 1;
      } else {
        finish_behind = true;
      }
      t->finish_pending = 0;
    }
  }
  cudaError_t le;
  if (t->early_reset && !actions && scan_was_valid) {
    // At most one CTA per SM, each walking its tiles: every CTA of the grid is then resident beside the predecessor's CTAs (a CTA that
    // has finished waits for the predecessor and keeps its slot, so CTAs beyond the resident set would only start when it drains).
    const int tile_ctas = d.num_tiles < t->sm_count ? d.num_tiles : t->sm_count;
    le = lm_launch(k_pre_early, dim3(tile_ctas + (grid > d.num_tiles ? 1 : 0)), dim3(T), 0, s, true, t->kp, a, tile_ctas);
  } else {
    le = lm_launch(k_pre, dim3(grid), dim3(T), 0, s, true, t->kp, a);
  }
  if (le != cudaSuccess) return lm_fail((int)le, "lm_task_pre_step: %s", cudaGetErrorString(le));
  if (finish_behind) {
    PostArgs f;
    memset(&f, 0, sizeof(f));
    f.buf = *buf; f.partials = t->partials; f.ticket = t->tickets + 1; f.tile_offsets = t->tile_offsets; f.scan_base = t->scan_base;
    f.tile_counts = t->tile_counts;
    KernelParams kpf = t->kp;
    kpf.d.scan_per = (d.num_tiles + FINISH_THREADS - 1) / FINISH_THREADS;
    k_post_finish<<<1, FINISH_THREADS, 0, s>>>(kpf, f);
  }
  return lm_check_launch("lm_task_pre_step");
}

int lm_task_reset_idx(LmTask* t, const LmRobotViews* loco, const LmRobotViews* mani, const float* env_pos,
                      const LmTaskBuffers* buf, const int64_t* env_ids, int32_t count, const float* uniforms, void* stream) {
  if (!t || !env_ids || !uniforms || count < 0) return lm_fail(LM_ERR_BAD_ARG, "lm_task_reset_idx: bad argument");
  int rc = check_views(t, loco, mani, env_pos, buf);
  if (rc) return rc;
  if (count == 0) return LM_OK;
  PreArgs a;
  memset(&a, 0, sizeof(a));
  if (loco) a.loco = *loco;
  if (mani) a.mani = *mani;
  a.env_pos = env_pos; a.buf = *buf; a.uniforms = uniforms; a.parity = t->parity ^ 1;   // "last" slot = current one before the next pre_step
  if (t->has_pd) a.pd = t->pd;
  k_reset_idx<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t->kp, a, env_ids, count);
  t->scan_valid = 0;   // reset_buf entries were cleared: the by-product scan of the last post_step is stale
  return lm_check_launch("lm_task_reset_idx");
}

int lm_task_post_step(LmTask* t, const LmRobotViews* loco, const LmRobotViews* mani, const float* env_pos,
                      const LmTaskBuffers* buf, void* stream) {
  if (!t) return lm_fail(LM_ERR_BAD_ARG, "lm_task_post_step: task is NULL");
  int rc = check_views(t, loco, mani, env_pos, buf);
  if (rc) return rc;
  PostArgs a;
  memset(&a, 0, sizeof(a));
  if (loco) a.loco = *loco;
  if (mani) a.mani = *mani;
  a.env_pos = env_pos; a.buf = *buf; a.partials = t->partials; a.ticket = t->tickets + 1; a.tile_offsets = t->tile_offsets;
  a.parity = t->parity; a.clamp_observations = t->clamp_observations; a.timeline = t->timeline;
  a.scan_base = t->scan_base; a.inline_finish = t->inline_finish; a.tile_counts = t->tile_counts;
  if (t->has_pd) {
    a.pd_targets = t->pd.targets; a.pd_last_targets = t->pd.last_targets; a.pd_target_obs_scale = t->pd.target_obs_scale;
    if (t->pd.carry_post_state) { a.pd_joint_pos = t->pd.joint_pos; a.pd_joint_vel = t->pd.joint_vel; }
  }
  if (t->kp.p.num_states == LM_NUM_STATES_FULL)
    k_post<<<t->kp.d.num_tiles, T, PostSmem::kBytes, (cudaStream_t)stream>>>(t->kp, a);
  else
  {
    cudaError_t le = a.timeline != nullptr
                         ? lm_launch(k_post_lean<true>, dim3(t->kp.d.num_tiles), dim3(T), LeanSmem::kBytes, (cudaStream_t)stream, true, t->kp, a)
                         : lm_launch(k_post_lean<false>, dim3(t->kp.d.num_tiles), dim3(T), LeanSmem::kBytes, (cudaStream_t)stream, true, t->kp, a);
    if (le != cudaSuccess) return lm_fail((int)le, "lm_task_post_step: %s", cudaGetErrorString(le));
  }
  if (!t->inline_finish && !t->defer_finish) k_post_finish<<<1, FINISH_THREADS, 0, (cudaStream_t)stream>>>(t->kp, a);
  // (a second post step before the pass of the first one ran overwrites its partial sums: that step's contribution to extras /
  //  counters is dropped, as documented for the deferred mode)
  if (t->defer_finish) t->finish_pending = 1;
  t->scan_valid = 1;
  return lm_check_launch("lm_task_post_step");
}

// Deferred finish pass: lm_task_post_step then only publishes the per-tile partial sums and reset counts — the post kernel sheds its
// serial tail (one CTA reading every tile's partials: 2.6 of 17.5 us at 65,536 envs).  The pass that turns them into extras /
// counters runs as ONE extra CTA of the next lm_task_pre_step's launch, next to the reset bodies (which take their rank base from
// the per-tile counts directly), or — when extras / counters are wanted before that, e.g. at a rollout boundary — through
// lm_task_post_finish.  Until one of the two has run, `extras`, `counters` and `num_reset` describe the step before.
int lm_task_set_defer_finish(LmTask* t, int on) {
  if (!t) return lm_fail(LM_ERR_BAD_ARG, "lm_task_set_defer_finish: task is NULL");
  if (t->finish_pending) return lm_fail(LM_ERR_BAD_ARG, "lm_task_set_defer_finish: a deferred finish pass is pending (lm_task_post_finish first)");
  t->defer_finish = on != 0;
  t->inline_finish = (t->kp.d.num_tiles <= INLINE_FINISH_MAX_TILES) && !t->defer_finish;
  // deferred: the pass runs with T threads inside k_pre (<= 1024 tiles); the stand-alone kernel gets its own scan_per at launch
  const int threads = (t->inline_finish || t->defer_finish) ? T : FINISH_THREADS;
  t->kp.d.scan_per = (t->kp.d.num_tiles + threads - 1) / threads;
  t->scan_valid = 0;                  // the layout of tile_offsets / scan_base depends on scan_per
  return LM_OK;
}

// Early reset (opt-in pipelining, with an action sink): lm_task_pre_step(actions == NULL) — the reset half only — no longer waits for
// the kernel launched before it on the stream; under programmatic dependent launch it runs beside that kernel and merely completes
// after it.  Contract: that kernel (a) does not write anything the reset half reads or writes — the policy forward with a sink
// writes the action ring's NEXT slot and the joint targets, the reset half the scene rows, goals, flags and the ring's LAST slot —
// and (b) triggers its dependents only after its own dependency wait, so that the post step whose flags the reset half reads has
// completed (lm_mlp_forward_sink does; any kernel without griddepcontrol instructions is trivially safe: no overlap then).
int lm_task_set_early_reset(LmTask* t, int on) {
  if (!t) return lm_fail(LM_ERR_BAD_ARG, "lm_task_set_early_reset: task is NULL");
  t->early_reset = on != 0;
  return LM_OK;
}

// Runs the deferred finish pass now (a no-op on the device when the last post step's pass already ran).
int lm_task_post_finish(LmTask* t, const LmTaskBuffers* buf, void* stream) {
  if (!t || !buf) return lm_fail(LM_ERR_BAD_ARG, "lm_task_post_finish: NULL argument");
  if (!t->defer_finish) return lm_fail(LM_ERR_BAD_ARG, "lm_task_post_finish: the task runs its finish pass itself (lm_task_set_defer_finish)");
  // (launched even when the host-side flag says nothing is pending: under CUDA-graph replay the flag does not follow the device,
  //  and the pass acts at most once per post step by itself)
  PostArgs a;
  memset(&a, 0, sizeof(a));
  a.buf = *buf; a.partials = t->partials; a.ticket = t->tickets + 1; a.tile_offsets = t->tile_offsets; a.scan_base = t->scan_base;
  a.tile_counts = t->tile_counts;
  KernelParams kpf = t->kp;
  kpf.d.scan_per = (t->kp.d.num_tiles + FINISH_THREADS - 1) / FINISH_THREADS;
  k_post_finish<<<1, FINISH_THREADS, 0, (cudaStream_t)stream>>>(kpf, a);
  t->finish_pending = 0;
  // the stand-alone pass left tile_offsets / scan_base in ITS layout: the next pre step must not mix them with the 128-thread one
  // (it takes its rank bases from the per-tile counts, which are layout-free)
  return lm_check_launch("lm_task_post_finish");
}

int lm_clamp_copy(const float* in, float* out, int64_t count, float clip, void* stream) {
  if (!in || !out || count < 0) return lm_fail(LM_ERR_BAD_ARG, "lm_clamp_copy: bad argument");
  if ((((uintptr_t)in) | ((uintptr_t)out)) & 15) return lm_fail(LM_ERR_ALIGNMENT, "lm_clamp_copy: unaligned");
  if (count == 0) return LM_OK;
  int blocks = (int)((count / 4 + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_clamp_copy<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, out, count, clip);
  return lm_check_launch("lm_clamp_copy");
}

static int step_host_impl(LmTask* t, const LmRobotViews* loco, const LmRobotViews* mani, const float* env_pos,
                          const LmTaskBuffers* buf, const LmHostStep* io, void* stream, bool sync) {
  if (!t || !io) return lm_fail(LM_ERR_BAD_ARG, "lm_task_step_host: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  // hs: where the physics-state copies run.  With a dedicated copy stream the shards' state copies queue back to back at full PCIe
  // bandwidth and post_physics_step on `s` waits for exactly the copy it reads; the small action / uniform copies stay on `s`
  // (they run under another shard's state copy).  Without one everything is in order on `s`.
  cudaStream_t hs = io->h2d_stream ? (cudaStream_t)io->h2d_stream : s;
  const bool split = hs != s;
  const int N = t->kp.p.num_envs;
  cudaError_t e;
  if (split && !t->have_events) {
    if ((e = cudaEventCreateWithFlags(&t->ev_state, cudaEventDisableTiming)) != cudaSuccess) goto fail;
    if ((e = cudaEventCreateWithFlags(&t->ev_consumed, cudaEventDisableTiming)) != cudaSuccess) goto fail;
    t->have_events = 1;
  }
  // pre_physics_step on the previous state, then the "physics" arrives from the host, then post_physics_step
  if ((e = cudaMemcpyAsync(io->actions_dev, io->actions_host, sizeof(float) * N * LM_NUM_ACT, cudaMemcpyHostToDevice, s)) != cudaSuccess) goto fail;
  if ((e = cudaMemcpyAsync(io->uniforms_dev, io->uniforms_host, sizeof(float) * N * 3, cudaMemcpyHostToDevice, s)) != cudaSuccess) goto fail;
  {
    int rc = lm_task_pre_step(t, loco, mani, env_pos, buf, io->actions_dev, io->uniforms_dev, stream);
    if (rc) return rc;
  }
  if (split) {
    // the state copy replaces the arena that (a) the previous post_physics_step read and (b) this pre_physics_step's reset scatter
    // wrote: both are behind this point of `s`
    if ((e = cudaEventRecord(t->ev_consumed, s)) != cudaSuccess) goto fail;
    if ((e = cudaStreamWaitEvent(hs, t->ev_consumed, 0)) != cudaSuccess) goto fail;
  }
  if (io->loco_state_host && io->loco_state_floats > 0)
    if ((e = cudaMemcpyAsync(io->loco_state_dev, io->loco_state_host, sizeof(float) * io->loco_state_floats, cudaMemcpyHostToDevice, hs)) != cudaSuccess) goto fail;
  if (io->mani_state_host && io->mani_state_floats > 0)
    if ((e = cudaMemcpyAsync(io->mani_state_dev, io->mani_state_host, sizeof(float) * io->mani_state_floats, cudaMemcpyHostToDevice, hs)) != cudaSuccess) goto fail;
  if (split) {
    if ((e = cudaEventRecord(t->ev_state, hs)) != cudaSuccess) goto fail;
    if ((e = cudaStreamWaitEvent(s, t->ev_state, 0)) != cudaSuccess) goto fail;
  }
  {
    int rc = lm_task_post_step(t, loco, mani, env_pos, buf, stream);
    if (rc) return rc;
  }
  if (io->obs_host && (e = cudaMemcpyAsync(io->obs_host, buf->obs_buf, sizeof(float) * (size_t)N * LM_NUM_OBS, cudaMemcpyDeviceToHost, s)) != cudaSuccess) goto fail;
  if (io->rew_host && (e = cudaMemcpyAsync(io->rew_host, buf->rew_buf, sizeof(float) * N, cudaMemcpyDeviceToHost, s)) != cudaSuccess) goto fail;
  if (io->reset_host && (e = cudaMemcpyAsync(io->reset_host, buf->reset_buf, sizeof(int64_t) * N, cudaMemcpyDeviceToHost, s)) != cudaSuccess) goto fail;
  if (sync && (e = cudaStreamSynchronize(s)) != cudaSuccess) goto fail;
  return LM_OK;
fail:
  return lm_fail((int)e, "lm_task_step_host: %s", cudaGetErrorString(e));
}

int lm_task_step_host(LmTask* t, const LmRobotViews* loco, const LmRobotViews* mani, const float* env_pos,
                      const LmTaskBuffers* buf, const LmHostStep* io, void* stream) {
  return step_host_impl(t, loco, mani, env_pos, buf, io, stream, true);
}

int lm_task_step_host_async(LmTask* t, const LmRobotViews* loco, const LmRobotViews* mani, const float* env_pos,
                            const LmTaskBuffers* buf, const LmHostStep* io, void* stream) {
  return step_host_impl(t, loco, mani, env_pos, buf, io, stream, false);
}

int lm_task_step_host_policy_async(LmTask* t, const LmRobotViews* loco, const LmRobotViews* mani, const float* env_pos,
                                   const LmTaskBuffers* buf, const LmHostStep* io, const LmHostPolicy* pol, void* stream) {
  if (!pol) return lm_fail(LM_ERR_BAD_ARG, "lm_task_step_host_policy_async: policy is NULL");
  int rc = step_host_impl(t, loco, mani, env_pos, buf, io, stream, false);
  if (rc) return rc;
  if (pol->kind == 0) return LM_OK;
  if (!pol->handle || !pol->mu_dev || !pol->value_dev) return lm_fail(LM_ERR_BAD_ARG, "lm_task_step_host_policy_async: policy handle / outputs missing");
  const int N = t->kp.p.num_envs;
  switch (pol->kind) {
    case 1: rc = lm_mlp_forward((LmMlp*)pol->handle, buf->obs_buf, pol->norm_mean, pol->norm_inv_std, pol->norm_clip, pol->mu_dev, pol->value_dev, N, stream); break;
    case 2: rc = lm_gnn_forward_bf16((LmGnn*)pol->handle, buf->obs_buf, pol->mu_dev, pol->value_dev, N, stream); break;
    case 3: rc = lm_gnn_forward((LmGnn*)pol->handle, buf->obs_buf, pol->mu_dev, pol->value_dev, N, stream); break;
    default: return lm_fail(LM_ERR_BAD_ARG, "lm_task_step_host_policy_async: unknown policy kind %d", pol->kind);
  }
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  if (pol->mu_host) e = cudaMemcpyAsync(pol->mu_host, pol->mu_dev, sizeof(float) * (size_t)N * LM_NUM_ACT, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && pol->value_host) e = cudaMemcpyAsync(pol->value_host, pol->value_dev, sizeof(float) * (size_t)N, cudaMemcpyDeviceToHost, s);
  return e == cudaSuccess ? LM_OK : lm_fail((int)e, "lm_task_step_host_policy_async: %s", cudaGetErrorString(e));
}

void* lm_host_alloc(int64_t bytes, int32_t write_combined) {
  void* p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, (size_t)bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
  if (e != cudaSuccess) { lm_set_error("lm_host_alloc: %s", cudaGetErrorString(e)); return nullptr; }
  return p;
}
void lm_host_free(void* p) { if (p) cudaFreeHost(p); }

int lm_stream_synchronize(void* stream) {
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  return e == cudaSuccess ? LM_OK : lm_fail((int)e, "lm_stream_synchronize: %s", cudaGetErrorString(e));
}

}  // extern "C"
