// PPO generalised advantage estimation as a reverse scan over a time-major [T,N] rollout.
// Not in the reference repo: rl_games 1.6.0 A2CBase.discount_values (pinned at
// /root/reference/RobotLearning/setup.py:20; used through cfg/train/*PPO.yaml:53-54,66) and skrl
// PPO compute_gae (scripts/skrl_ppo_locomanipulation.py:104-105).  Both store rollouts time-major,
// so one thread owns 4 adjacent envs (128-bit coalesced loads) and walks t = T-1 .. 0 with the
// loads of a 4-step chunk issued before the dependent arithmetic.  A second kernel handles an
// env-major [N,T] buffer with a warp-level affine scan (32 time steps per shuffle scan).
#include "lm_common.cuh"
#include "lm_host.h"

namespace lm {

struct F4 { float v[4]; };
__device__ __forceinline__ F4 ld4(const float* p) { float4 x = *reinterpret_cast<const float4*>(p); return F4{{x.x, x.y, x.z, x.w}}; }
__device__ __forceinline__ void st4(float* p, const F4& a) { *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }

// one GAE step, op order of the two libraries (see oracle/port_policy.py)
template <int kVariant>
__device__ __forceinline__ float gae_step(float r, float v, float d_here, float d_next, float v_next, float adv, float gamma, float gl,
                                          float lambda) {
  if (kVariant == LM_GAE_RLGAMES) {
    // delta = r + gamma * v_next * nnt - v ; adv = delta + (gamma*tau) * nnt * adv       (next-step done mask)
    const float nnt = sub(1.0f, d_next);
    const float delta = sub(add(r, mul(mul(gamma, v_next), nnt)), v);
    return add(delta, mul(mul(gl, nnt), adv));
  } else {
    // adv = r - v + gamma * not_done * (v_next + lambda * adv)                            (current-step done mask)
    const float nd = d_here != 0.0f ? 0.0f : 1.0f;
    return add(sub(r, v), mul(mul(gamma, nd), add(v_next, mul(lambda, adv))));
  }
}

template <int kVariant, int kVec>
__global__ void __launch_bounds__(128) k_gae_time_major(const float* __restrict__ rewards, const float* __restrict__ values,
                                                        const float* __restrict__ dones, const float* __restrict__ last_values,
                                                        const float* __restrict__ last_dones, float* __restrict__ adv_out,
                                                        float* __restrict__ ret_out, int T, int N, float gamma, float gl, float lambda) {
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * kVec;
  if (col >= N) return;
  float adv[kVec], v_next[kVec], d_next[kVec];
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    adv[k] = 0.0f;
    v_next[k] = last_values[col + k];
    d_next[k] = last_dones ? last_dones[col + k] : 0.0f;
  }
  constexpr int CH = 4;
  for (int t1 = T; t1 > 0; t1 -= CH) {
    const int t0 = max(0, t1 - CH);
    float r[CH][kVec], v[CH][kVec], d[CH][kVec];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int t = t1 - 1 - c;
      if (t >= t0) {
        const size_t off = (size_t)t * N + col;
        if (kVec == 4) {
          F4 a = ld4(rewards + off), b = ld4(values + off), e = ld4(dones + off);
#pragma unroll
          for (int k = 0; k < kVec; ++k) { r[c][k] = a.v[k]; v[c][k] = b.v[k]; d[c][k] = e.v[k]; }
        } else {
#pragma unroll
          for (int k = 0; k < kVec; ++k) { r[c][k] = rewards[off + k]; v[c][k] = values[off + k]; d[c][k] = dones[off + k]; }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int t = t1 - 1 - c;
      if (t >= t0) {
        F4 a, rt;
#pragma unroll
        for (int k = 0; k < kVec; ++k) {
          adv[k] = gae_step<kVariant>(r[c][k], v[c][k], d[c][k], d_next[k], v_next[k], adv[k], gamma, gl, lambda);
          a.v[k] = adv[k];
          rt.v[k] = add(adv[k], v[c][k]);
          v_next[k] = v[c][k];
          d_next[k] = d[c][k];
        }
        const size_t off = (size_t)t * N + col;
        if (kVec == 4) { st4(adv_out + off, a); if (ret_out) st4(ret_out + off, rt); }
        else {
#pragma unroll
          for (int k = 0; k < kVec; ++k) { adv_out[off + k] = a.v[k]; if (ret_out) ret_out[off + k] = rt.v[k]; }
        }
      }
    }
  }
}

// Env-major [N,T] rollouts: A_t = delta_t + c_t * A_{t+1} is an affine recurrence; compose (c, delta)
// pairs with a warp shuffle scan over 32 reversed time steps at a time (one warp per env).
// Association differs from the sequential loop, so results agree to fp32 round-off (1e-6 relative).
template <int kVariant>
__global__ void __launch_bounds__(128) k_gae_env_major(const float* __restrict__ rewards, const float* __restrict__ values,
                                                       const float* __restrict__ dones, const float* __restrict__ last_values,
                                                       const float* __restrict__ last_dones, float* __restrict__ adv_out,
                                                       float* __restrict__ ret_out, int T, int N, float gamma, float gl, float lambda) {
  const int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (env >= N) return;
  const size_t base = (size_t)env * T;
  float carry = 0.0f;   // A_{t+1} entering the current 32-step window
  for (int hi = T; hi > 0; hi -= 32) {
    const int t = hi - 1 - lane;   // lane 0 = latest step of the window
    float c = 0.0f, dl = 0.0f, v = 0.0f;
    if (t >= 0) {
      v = values[base + t];
      const float r = rewards[base + t];
      const float v_next = (t == T - 1) ? last_values[env] : values[base + t + 1];
      if (kVariant == LM_GAE_RLGAMES) {
        const float d_next = (t == T - 1) ? (last_dones ? last_dones[env] : 0.0f) : dones[base + t + 1];
        const float nnt = 1.0f - d_next;
        dl = r + gamma * v_next * nnt - v;
        c = gl * nnt;
      } else {
        const float nd = dones[base + t] != 0.0f ? 0.0f : 1.0f;
        dl = r - v + gamma * nd * v_next;
        c = gamma * nd * lambda;
      }
    }
    // inclusive scan of affine maps x -> dl + c x, composing lane i with lanes < i (later time steps)
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float c_up = __shfl_up_sync(0xffffffffu, c, off);
      const float d_up = __shfl_up_sync(0xffffffffu, dl, off);
      if (lane >= off) { dl = dl + c * d_up; c = c * c_up; }
    }
    const float a = dl + c * carry;
    if (t >= 0) { adv_out[base + t] = a; if (ret_out) ret_out[base + t] = a + v; }
    carry = __shfl_sync(0xffffffffu, a, 31);
  }
}

}  // namespace lm

using namespace lm;

template <int V>
static void launch_tm(const float* r, const float* v, const float* d, const float* lv, const float* ld, float* a, float* ret, int T,
                      int N, float g, float gl, float lam, cudaStream_t s) {
  const bool vec = (N % 4 == 0) && !((((uintptr_t)r) | ((uintptr_t)v) | ((uintptr_t)d) | ((uintptr_t)a) | ((uintptr_t)ret)) & 15);
  if (vec) k_gae_time_major<V, 4><<<(N / 4 + 127) / 128, 128, 0, s>>>(r, v, d, lv, ld, a, ret, T, N, g, gl, lam);
  else k_gae_time_major<V, 1><<<(N + 127) / 128, 128, 0, s>>>(r, v, d, lv, ld, a, ret, T, N, g, gl, lam);
}

extern "C" {

int lm_gae(const float* rewards, const float* values, const float* dones, const float* last_values, const float* last_dones,
           float* advantages, float* returns, int32_t T, int32_t N, float gamma, float lambda, int32_t variant, void* stream) {
  if (!rewards || !values || !dones || !last_values || !advantages || T <= 0 || N <= 0)
    return lm_fail(LM_ERR_BAD_ARG, "lm_gae: bad argument");
  const float gl = (float)((double)gamma * (double)lambda);   // Python evaluates gamma * tau in double first
  cudaStream_t s = (cudaStream_t)stream;
  if (variant == LM_GAE_RLGAMES) launch_tm<LM_GAE_RLGAMES>(rewards, values, dones, last_values, last_dones, advantages, returns, T, N, gamma, gl, lambda, s);
  else if (variant == LM_GAE_SKRL) launch_tm<LM_GAE_SKRL>(rewards, values, dones, last_values, last_dones, advantages, returns, T, N, gamma, gl, lambda, s);
  else return lm_fail(LM_ERR_UNSUPPORTED, "lm_gae: unknown variant %d", variant);
  return lm_check_launch("lm_gae");
}

int lm_gae_env_major(const float* rewards, const float* values, const float* dones, const float* last_values, const float* last_dones,
                     float* advantages, float* returns, int32_t T, int32_t N, float gamma, float lambda, int32_t variant, void* stream) {
  if (!rewards || !values || !dones || !last_values || !advantages || T <= 0 || N <= 0)
    return lm_fail(LM_ERR_BAD_ARG, "lm_gae_env_major: bad argument");
  const float gl = (float)((double)gamma * (double)lambda);
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = (int)(((int64_t)N * 32 + 127) / 128);
  if (variant == LM_GAE_RLGAMES) k_gae_env_major<LM_GAE_RLGAMES><<<blocks, 128, 0, s>>>(rewards, values, dones, last_values, last_dones, advantages, returns, T, N, gamma, gl, lambda);
  else if (variant == LM_GAE_SKRL) k_gae_env_major<LM_GAE_SKRL><<<blocks, 128, 0, s>>>(rewards, values, dones, last_values, last_dones, advantages, returns, T, N, gamma, gl, lambda);
  else return lm_fail(LM_ERR_UNSUPPORTED, "lm_gae_env_major: unknown variant %d", variant);
  return lm_check_launch("lm_gae_env_major");
}

}  // extern "C"
