// Explicit position-PD torque loop of the *_position_control task variants:
//   quadruped_pose_control_position_control.py:275-313, quadruped_manipulate_plate_position_control.py:276-316,
//   joint_locomanipulation_position_control.py:349-408.
// Two elementwise kernels over [n,12]: the swing/extension target integration done once per env step and
// the torque law run once per physics sub-step (PhysX steps in between, so they cannot be fused).
#include "lm_common.cuh"
#include "lm_host.h"

namespace lm {

struct PdIndex { int32_t idx[LM_NUM_ACT]; };
struct PdLimits { float lo[LM_NUM_ACT], hi[LM_NUM_ACT]; };

// torque = clamp(kp (target - q) - kd qd, ±max); torque_sum += torque          (…position_control.py:295-311)
__global__ void __launch_bounds__(256) k_pd_substep(const float* __restrict__ dof_pos, const float* __restrict__ dof_vel,
                                                    const float* __restrict__ targets, float* __restrict__ torque,
                                                    float* __restrict__ torque_sum, PdIndex ix, int n, float kp, float kd, float max_effort) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (env, 4 joints)
  if (i >= n * 3) return;
  const int env = i / 3, j0 = (i - env * 3) * 4;
  const float4 tg = reinterpret_cast<const float4*>(targets)[i];
  float4 ts = torque_sum ? reinterpret_cast<float4*>(torque_sum)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float tgt[4] = {tg.x, tg.y, tg.z, tg.w};
  float out[4], acc[4] = {ts.x, ts.y, ts.z, ts.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float q = dof_pos[(size_t)env * LM_NUM_DOF + ix.idx[j0 + k]];
    const float qd = dof_vel[(size_t)env * LM_NUM_DOF + ix.idx[j0 + k]];
    float tau = sub(mul(kp, sub(tgt[k], q)), mul(kd, qd));
    tau = fminf(fmaxf(tau, -max_effort), max_effort);
    out[k] = tau;
    acc[k] = add(acc[k], tau);
  }
  reinterpret_cast<float4*>(torque)[i] = make_float4(out[0], out[1], out[2], out[3]);
  if (torque_sum) reinterpret_cast<float4*>(torque_sum)[i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// update_joint_states (robot.py:283-291) + torque law on the refreshed state.  One thread per (env, 4 joints).
__global__ void __launch_bounds__(256) k_pd_update_torque(const float* __restrict__ dof_pos, const float* __restrict__ dof_vel, PdIndex ix,
                                                          float* __restrict__ joint_pos, float* __restrict__ joint_vel,
                                                          float* __restrict__ last_joint_vel, const float* __restrict__ targets,
                                                          float* __restrict__ torque, float* __restrict__ torque_sum,
                                                          float* __restrict__ mean_torque, float divisor, int n, float kp, float kd,
                                                          float max_effort, int gather, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  const int env = i / 3, j0 = (i - env * 3) * 4;
  float q[4], qd[4];
  if (gather) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      q[k] = dof_pos[(size_t)env * LM_NUM_DOF + ix.idx[j0 + k]];
      qd[k] = dof_vel[(size_t)env * LM_NUM_DOF + ix.idx[j0 + k]];
    }
    reinterpret_cast<float4*>(joint_pos)[i] = make_float4(q[0], q[1], q[2], q[3]);
    reinterpret_cast<float4*>(joint_vel)[i] = make_float4(qd[0], qd[1], qd[2], qd[3]);
    reinterpret_cast<float4*>(last_joint_vel)[i] = make_float4(qd[0], qd[1], qd[2], qd[3]);     // robot.py:291
  } else {
    const float4 a = reinterpret_cast<const float4*>(joint_pos)[i], b = reinterpret_cast<const float4*>(joint_vel)[i];
    q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; qd[0] = b.x; qd[1] = b.y; qd[2] = b.z; qd[3] = b.w;
  }
  const float4 tg = reinterpret_cast<const float4*>(targets)[i];
  const float4 ts = reinterpret_cast<const float4*>(torque_sum)[i];
  const float tgt[4] = {tg.x, tg.y, tg.z, tg.w};
  float out[4], acc[4] = {ts.x, ts.y, ts.z, ts.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float tau = sub(mul(kp, sub(tgt[k], q[k])), mul(kd, qd[k]));
    tau = fminf(fmaxf(tau, -max_effort), max_effort);
    out[k] = tau;
    if (accumulate) acc[k] = add(acc[k], tau);
  }
  reinterpret_cast<float4*>(torque)[i] = make_float4(out[0], out[1], out[2], out[3]);
  if (accumulate) reinterpret_cast<float4*>(torque_sum)[i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  if (mean_torque)
    reinterpret_cast<float4*>(mean_torque)[i] = make_float4(fdiv(acc[0], divisor), fdiv(acc[1], divisor), fdiv(acc[2], divisor), fdiv(acc[3], divisor));
}

// se += action_scale * a; clamp to swing/extension limits; dof1 = se[0:4], dof2 = swing + ext/2,
// dof3 = swing - ext/2                                                              (…position_control.py:275-287)
__global__ void __launch_bounds__(128) k_pd_integrate(const float* __restrict__ actions, float* __restrict__ targets_se,
                                                      float* __restrict__ targets, PdLimits lim, int n, float action_scale, int zero_actions) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  float se[LM_NUM_ACT];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 a = reinterpret_cast<const float4*>(actions)[env * 3 + k];
    const float4 s = reinterpret_cast<const float4*>(targets_se)[env * 3 + k];
    const float av[4] = {a.x, a.y, a.z, a.w}, sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = 4 * k + q;
      const float delta = mul(zero_actions ? 0.0f : av[q], action_scale);
      se[j] = fminf(fmaxf(add(sv[q], delta), lim.lo[j]), lim.hi[j]);
    }
    reinterpret_cast<float4*>(targets_se)[env * 3 + k] = make_float4(se[4 * k], se[4 * k + 1], se[4 * k + 2], se[4 * k + 3]);
  }
  float tg[LM_NUM_ACT];
#pragma unroll
  for (int j = 0; j < 4; ++j) tg[j] = se[j];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float swing = se[4 + 2 * k], ext = se[5 + 2 * k];
    tg[4 + 2 * k] = add(swing, fdiv(ext, 2.0f));
    tg[5 + 2 * k] = sub(swing, fdiv(ext, 2.0f));
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
    reinterpret_cast<float4*>(targets)[env * 3 + k] = make_float4(tg[4 * k], tg[4 * k + 1], tg[4 * k + 2], tg[4 * k + 3]);
}

}  // namespace lm

using namespace lm;

extern "C" {

int lm_pd_substep(const float* dof_pos, const float* dof_vel, const float* targets, float* torque, float* torque_sum,
                  const int32_t* dof_index_host, int32_t n, float kp, float kd, float max_effort, void* stream) {
  if (!dof_pos || !dof_vel || !targets || !torque || n <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_pd_substep: bad argument");
  if ((((uintptr_t)targets) | ((uintptr_t)torque) | ((uintptr_t)torque_sum)) & 15) return lm_fail(LM_ERR_ALIGNMENT, "lm_pd_substep: unaligned");
  PdIndex ix;
  for (int j = 0; j < LM_NUM_ACT; ++j) {
    ix.idx[j] = dof_index_host ? dof_index_host[j] : j;
    if (ix.idx[j] < 0 || ix.idx[j] >= LM_NUM_DOF) return lm_fail(LM_ERR_BAD_ARG, "lm_pd_substep: dof index out of range");
  }
  k_pd_substep<<<(n * 3 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dof_pos, dof_vel, targets, torque, torque_sum, ix, n, kp, kd, max_effort);
  return lm_check_launch("lm_pd_substep");
}

int lm_pd_torque(const float* joint_pos, const float* joint_vel, const float* targets, float* torque, float* torque_sum, int32_t n,
                 float kp, float kd, float max_effort, void* stream) {
  if (!joint_pos || !joint_vel || !targets || !torque || !torque_sum || n <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_pd_torque: bad argument");
  if ((((uintptr_t)joint_pos) | ((uintptr_t)joint_vel) | ((uintptr_t)targets) | ((uintptr_t)torque) | ((uintptr_t)torque_sum)) & 15)
    return lm_fail(LM_ERR_ALIGNMENT, "lm_pd_torque: unaligned");
  PdIndex ix;
  for (int j = 0; j < LM_NUM_ACT; ++j) ix.idx[j] = j;
  k_pd_update_torque<<<(n * 3 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(nullptr, nullptr, ix, const_cast<float*>(joint_pos),
                                                                          const_cast<float*>(joint_vel), nullptr, targets, torque, torque_sum,
                                                                          nullptr, 1.0f, n, kp, kd, max_effort, 0, 1);
  return lm_check_launch("lm_pd_torque");
}

int lm_pd_update_and_torque(const float* dof_pos, const float* dof_vel, const int32_t* dof_index_host, float* joint_pos, float* joint_vel,
                            float* last_joint_vel, const float* targets, float* torque, float* torque_sum, int32_t accumulate,
                            float* mean_torque, float torque_divisor, int32_t n, float kp, float kd, float max_effort, void* stream) {
  if (!dof_pos || !dof_vel || !joint_pos || !joint_vel || !last_joint_vel || !targets || !torque || !torque_sum || n <= 0)
    return lm_fail(LM_ERR_BAD_ARG, "lm_pd_update_and_torque: bad argument");
  if ((((uintptr_t)joint_pos) | ((uintptr_t)joint_vel) | ((uintptr_t)last_joint_vel) | ((uintptr_t)targets) | ((uintptr_t)torque) |
       ((uintptr_t)torque_sum) | ((uintptr_t)mean_torque)) & 15)
    return lm_fail(LM_ERR_ALIGNMENT, "lm_pd_update_and_torque: unaligned");
  if (mean_torque && torque_divisor == 0.0f) return lm_fail(LM_ERR_BAD_ARG, "lm_pd_update_and_torque: zero divisor");
  PdIndex ix;
  for (int j = 0; j < LM_NUM_ACT; ++j) {
    ix.idx[j] = dof_index_host ? dof_index_host[j] : j;
    if (ix.idx[j] < 0 || ix.idx[j] >= LM_NUM_DOF) return lm_fail(LM_ERR_BAD_ARG, "lm_pd_update_and_torque: dof index out of range");
  }
  k_pd_update_torque<<<(n * 3 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dof_pos, dof_vel, ix, joint_pos, joint_vel, last_joint_vel, targets,
                                                                          torque, torque_sum, mean_torque, torque_divisor, n, kp, kd, max_effort, 1, accumulate);
  return lm_check_launch("lm_pd_update_and_torque");
}

int lm_pd_integrate_targets(const float* actions, float* targets_se, float* targets, const float* se_lower_host,
                            const float* se_upper_host, int32_t n, float action_scale, int32_t zero_actions, void* stream) {
  if (!actions || !targets_se || !targets || !se_lower_host || !se_upper_host || n <= 0)
    return lm_fail(LM_ERR_BAD_ARG, "lm_pd_integrate_targets: bad argument");
  if ((((uintptr_t)actions) | ((uintptr_t)targets_se) | ((uintptr_t)targets)) & 15) return lm_fail(LM_ERR_ALIGNMENT, "lm_pd_integrate_targets: unaligned");
  PdLimits lim;
  for (int j = 0; j < LM_NUM_ACT; ++j) { lim.lo[j] = se_lower_host[j]; lim.hi[j] = se_upper_host[j]; }
  k_pd_integrate<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(actions, targets_se, targets, lim, n, action_scale, zero_actions);
  return lm_check_launch("lm_pd_integrate_targets");
}

}  // extern "C"
