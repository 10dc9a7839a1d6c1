// *_custom_controller.py task variants: the rows the variant adds on top of the shared task kernels.
// Reference: tasks/quadruped_pose_control_tasks/quadruped_pose_control_custom_controller.py (cited per line below) and its
// manipulation sibling tasks/quadruped_manipulate_plate/quadruped_manipulate_plate_custom_controller.py (same rows, minus the
// rot-dist-decrease term and the per-step refresh of last_joint_position_targets).
// k_pre / k_post (lm_task.cu) run unchanged; k_cc_pre / k_cc_post are one thread per env, straight-line.
#include "lm_common.cuh"
#include "lm_host.h"
#include "lm_reduce.cuh"

void lm_rollout_workspace(LmRollout* r, double** partials, unsigned int** ticket, int* sm_count);   // lm_rollout.cu (library-internal)

namespace lm {
namespace cc {

constexpr int THREADS = 256;
constexpr int NRED = 6;          // sums: action rate, mechanical power, target error, rot-dist decrease, NaN count, < -50 count

// 2 asin(min(|q_xyz|, 1)) exactly as k_post evaluates it (lm_task.cu pose_main): torch.norm + asin rounded once from double
__device__ __forceinline__ float rot_distance(const Quat& d) {
  return mul(2.0f, (float)asin((double)fminf(norm3_torch(d.x, d.y, d.z), 1.0f)));
}

__global__ void k_cc_pre(const LmCustomController c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.num_reset[0]) return;
  const int e = (int)c.reset_env_ids[i];
  const float4 g = *reinterpret_cast<const float4*>(c.goal + (size_t)e * 4);
  const Quat qd{c.default_quaternion[0], c.default_quaternion[1], c.default_quaternion[2], c.default_quaternion[3]};
  // :384-386  quat_mul(conj(default quaternion), conj(goal))
  c.last_rot_dist[e] = rot_distance(quat_mul(quat_conj(qd), quat_conj(Quat{g.x, g.y, g.z, g.w})));
}

struct PostArgs {
  LmCustomController c;
  double* partials;
  unsigned int* ticket;
};

__global__ void __launch_bounds__(THREADS) k_cc_post(const PostArgs a) {
  __shared__ double sm[THREADS / 32][NRED];
  const LmCustomController& c = a.c;
  const int tid = threadIdx.x;
  const int e = blockIdx.x * THREADS + tid;
  float red[NRED] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (e < c.num_envs) {
    // ---- raw orientation error (:430, :515-516).  The frame quaternion (ground_quaternion = conj(base quaternion); plate quaternion
    // in the robot frame for the manipulation file) is what k_post left in the states row, columns 6..9 (quaternion scale 1).
    float* st = c.states + (size_t)e * 93;
    const float4 g4 = *reinterpret_cast<const float4*>(c.goal + (size_t)e * 4);
    const Quat diff = quat_mul(Quat{st[6], st[7], st[8], st[9]}, quat_conj(Quat{g4.x, g4.y, g4.z, g4.w}));
    const bool rot_term = (c.flags & LM_CC_ROT_DIST_TERM) != 0;
    const float rot_dist = rot_term ? rot_distance(diff) : 0.0f;
    const float lim = c.clip_observations;
    float dq[4] = {diff.w, diff.x, diff.y, diff.z};
    if (c.clamp_observations) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dq[k] = fminf(fmaxf(dq[k], -lim), lim);
    }
    // ---- joint state after the step (update_all_states), targets, torque
    float jp[12], jv[12], tg[12], ltg[12], tq[12], cur[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      jp[j] = c.dof_pos[(size_t)e * 20 + c.dof_index[j]];
      jv[j] = c.dof_vel[(size_t)e * 20 + c.dof_index[j]];
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float4 t4 = reinterpret_cast<const float4*>(c.targets + (size_t)e * 12)[q];
      const float4 l4 = reinterpret_cast<const float4*>(c.last_targets + (size_t)e * 12)[q];
      const float4 m4 = reinterpret_cast<const float4*>(c.torque + (size_t)e * 12)[q];
      tg[4 * q] = t4.x; tg[4 * q + 1] = t4.y; tg[4 * q + 2] = t4.z; tg[4 * q + 3] = t4.w;
      ltg[4 * q] = l4.x; ltg[4 * q + 1] = l4.y; ltg[4 * q + 2] = l4.z; ltg[4 * q + 3] = l4.w;
      tq[4 * q] = m4.x; tq[4 * q + 1] = m4.y; tq[4 * q + 2] = m4.z; tq[4 * q + 3] = m4.w;
    }
    // ---- observation row (:441-455): scratch row with the raw quat_diff, then the two target rows
    const float4* o64 = reinterpret_cast<const float4*>(c.obs64 + (size_t)e * 64);
    float4* o88 = reinterpret_cast<float4*>(c.obs + (size_t)e * 88);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float4 v = o64[k];
      if (k == 1) { v.z = dq[0]; v.w = dq[1]; }          // obs[6:10] = quat_diff (w, x, y, z)
      if (k == 2) { v.x = dq[2]; v.y = dq[3]; }
      if (k >= 10 && k < 13) { cur[4 * (k - 10)] = v.x; cur[4 * (k - 10) + 1] = v.y; cur[4 * (k - 10) + 2] = v.z; cur[4 * (k - 10) + 3] = v.w; }
      o88[k] = v;
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      float4 t = make_float4(mul(tg[4 * q], c.target_obs_scale), mul(tg[4 * q + 1], c.target_obs_scale), mul(tg[4 * q + 2], c.target_obs_scale),
                             mul(tg[4 * q + 3], c.target_obs_scale));
      float4 l = make_float4(mul(ltg[4 * q], c.target_obs_scale), mul(ltg[4 * q + 1], c.target_obs_scale), mul(ltg[4 * q + 2], c.target_obs_scale),
                             mul(ltg[4 * q + 3], c.target_obs_scale));
      if (c.clamp_observations) {
        t.x = fminf(fmaxf(t.x, -lim), lim); t.y = fminf(fmaxf(t.y, -lim), lim); t.z = fminf(fmaxf(t.z, -lim), lim); t.w = fminf(fmaxf(t.w, -lim), lim);
        l.x = fminf(fmaxf(l.x, -lim), lim); l.y = fminf(fmaxf(l.y, -lim), lim); l.z = fminf(fmaxf(l.z, -lim), lim); l.w = fminf(fmaxf(l.w, -lim), lim);
      }
      o88[16 + q] = t;
      o88[19 + q] = l;
    }
    // ---- states row: quat_diff columns 41..44 (:474)
#pragma unroll
    for (int k = 0; k < 4; ++k) st[41 + k] = dq[k];
    // ---- reward terms (:531-546)
    float a_sum = 0.0f, mech = 0.0f, terr = 0.0f;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      a_sum = add(a_sum, fabsf(cur[j]));
      mech = add(mech, fabsf(mul(tq[j], jv[j])));
      terr = add(terr, fabsf(sub(ltg[j], jp[j])));
    }
    const float arate = mul(a_sum, c.action_rate_scale);
    const float mech_pen = mul(c.mechanical_power_penalty_scale, mech);
    const float terr_pen = mul(c.position_target_error_penalty_scale, terr);
    float rdec = 0.0f;
    float total = add(add(add(c.rew[e], arate), mech_pen), terr_pen);
    if (rot_term) {
      const float larger = rot_dist > c.no_rot_dist_decreasing_reward_thresh ? 1.0f : 0.0f;
      rdec = mul(mul(larger, sub(c.last_rot_dist[e], rot_dist)), c.rot_dist_decreasing_reward_scale);
      c.last_rot_dist[e] = rot_dist;
      total = add(total, rdec);
    }
    if (c.clip_reward) total = fmaxf(total, 0.0f);                                  // :619-620 (fall_penalty is 0 in the file)
    c.rew[e] = total;
    red[0] = arate; red[1] = mech_pen; red[2] = terr_pen; red[3] = rdec;
    red[4] = isnan(total) ? 1.0f : 0.0f; red[5] = total < -50.0f ? 1.0f : 0.0f;
    // ---- bookkeeping: last_joint_position_targets (:724), carried joint state (:392-393)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      if (c.flags & LM_CC_UPDATE_LAST_TARGETS)
        reinterpret_cast<float4*>(c.last_targets + (size_t)e * 12)[q] = make_float4(tg[4 * q], tg[4 * q + 1], tg[4 * q + 2], tg[4 * q + 3]);
      reinterpret_cast<float4*>(c.joint_pos + (size_t)e * 12)[q] = make_float4(jp[4 * q], jp[4 * q + 1], jp[4 * q + 2], jp[4 * q + 3]);
      reinterpret_cast<float4*>(c.joint_vel + (size_t)e * 12)[q] = make_float4(jv[4 * q], jv[4 * q + 1], jv[4 * q + 2], jv[4 * q + 3]);
    }
  }
  // ---- means of the new terms (extras), deterministic: per-CTA doubles -> last CTA
  {
    const double d = cta_sum<NRED, THREADS>(red, sm);
    if (tid < NRED) a.partials[(size_t)blockIdx.x * NRED + tid] = d;
  }
  if (!last_cta(a.ticket)) return;
  {
    __shared__ double part[32][NRED];
    const double d = grid_sum<NRED, 32>(a.partials, part);
    if (tid < NRED) {
      const double n = (double)c.num_envs;
      if (tid == 0) c.extras[LM_EXTRA_ACTION_RATE_PENALTY] = (float)(d / n);
      if (tid == 1) c.extras[LM_EXTRA_MECHANICAL_POWER_PENALTY] = (float)(d / n);
      if (tid == 2) c.extras[LM_EXTRA_TARGET_ERROR_PENALTY] = (float)(d / n);
      if (tid == 3 && (c.flags & LM_CC_ROT_DIST_TERM)) c.extras[LM_EXTRA_ROT_DIST_DECREASING_REW] = (float)(d / n);
      if (tid == 4) c.extras[LM_EXTRA_NAN_COUNT] = (float)d;
      if (tid == 5) c.extras[LM_EXTRA_BELOW_M50_COUNT] = (float)d;
    }
  }
}

}  // namespace cc
}  // namespace lm

using namespace lm;

static int cc_check(const LmCustomController* c, const char* who, bool post) {
  if (!c) return lm_fail(LM_ERR_BAD_ARG, "%s: NULL", who);
  if (c->num_envs <= 0 || !c->goal) return lm_fail(LM_ERR_BAD_ARG, "%s: bad num_envs / NULL tensor", who);
  if ((c->flags & LM_CC_ROT_DIST_TERM) && !c->last_rot_dist) return lm_fail(LM_ERR_BAD_ARG, "%s: last_rot_dist is NULL", who);
  if (!post && !(c->flags & LM_CC_ROT_DIST_TERM)) return lm_fail(LM_ERR_BAD_ARG, "%s: nothing to do without the rot-dist term", who);
  if (!post) return (c->reset_env_ids && c->num_reset) ? LM_OK : lm_fail(LM_ERR_BAD_ARG, "%s: reset id list missing", who);
  const void* ptrs[] = {c->obs64, c->dof_pos, c->dof_vel, c->torque, c->targets, c->last_targets, c->joint_pos, c->joint_vel,
                        c->obs, c->states, c->rew, c->extras};
  for (const void* p : ptrs) if (!p) return lm_fail(LM_ERR_BAD_ARG, "%s: NULL tensor", who);
  const void* al[] = {c->obs64, c->goal, c->torque, c->targets, c->last_targets, c->joint_pos, c->joint_vel, c->obs};
  for (const void* p : al) if (((uintptr_t)p) & 15) return lm_fail(LM_ERR_ALIGNMENT, "%s: tensor not 16-byte aligned", who);
  return LM_OK;
}

extern "C" {

int lm_cc_pre(const LmCustomController* c, void* stream) {
  const int rc = cc_check(c, "lm_cc_pre", false);
  if (rc != LM_OK) return rc;
  cc::k_cc_pre<<<(c->num_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*c);
  return lm_check_launch("lm_cc_pre");
}

int lm_cc_post(LmRollout* workspace, const LmCustomController* c, void* stream) {
  if (!workspace) return lm_fail(LM_ERR_BAD_ARG, "lm_cc_post: NULL workspace");
  const int rc = cc_check(c, "lm_cc_post", true);
  if (rc != LM_OK) return rc;
  cc::PostArgs a;
  a.c = *c;
  int sm_count = 0;
  lm_rollout_workspace(workspace, &a.partials, &a.ticket, &sm_count);
  const int grid = (c->num_envs + cc::THREADS - 1) / cc::THREADS;
  if ((long long)grid * cc::NRED > (long long)sm_count * 128) return lm_fail(LM_ERR_UNSUPPORTED, "lm_cc_post: more than %d envs per call", sm_count * 128 / cc::NRED * cc::THREADS);
  cc::k_cc_post<<<grid, cc::THREADS, 0, (cudaStream_t)stream>>>(a);
  return lm_check_launch("lm_cc_post");
}

}  // extern "C"
