// PPO update pieces: loss head + fused clip / Adam step.
//
// Not in the reference repo.  rl_games 1.6.0 (RobotLearning/setup.py:20) a2c_continuous.calc_gradients with
// common_losses.actor_loss / critic_loss, bound_loss, policy_kl, configured by cfg/train/JointLocomanipulationPPO.yaml:52-72;
// restated in oracle/port_policy.py (ppo_loss) where torch autograd supplies the reference gradients.
//
// k_ppo_loss: one thread per sample, 128-bit row accesses, 13 reductions (5 loss terms + clip fraction, 12 log-std
// gradients ... ) through warp shuffles -> per-CTA doubles -> last CTA.  About 270 B read and 52 B written per sample:
// HBM-bound, one launch instead of the ~40 element-wise launches of the torch graph (forward + backward).
#include "lm_common.cuh"
#include "lm_host.h"
#include "lm_reduce.cuh"

namespace lm {
namespace ppo {

constexpr int THREADS = 256;
constexpr int NRED = 20;     // a_loss, c_loss, b_loss, kl, clipped, 12 x dlogstd, 3 spare

struct Args {
  LmPpoBatch b;
  float* grad_mu; float* grad_value; float* grad_log_std; float* stats;
  double* partials; unsigned int* ticket;
};

__device__ __forceinline__ void ld12(const float* p, float v[12]) {
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(p) + q);
    v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
  }
}

__global__ void __launch_bounds__(THREADS) k_ppo_loss(const Args a) {
  __shared__ double sm[THREADS / 32][NRED];
  __shared__ float s_sigma[12], s_inv_var[12], s_old_sigma[12];
  const LmPpoBatch& B = a.b;
  const int tid = threadIdx.x;
  if (tid < 12) {
    const float sg = expf(B.log_std[tid]);
    s_sigma[tid] = sg;
    s_inv_var[tid] = 1.0f / (sg * sg);
    s_old_sigma[tid] = B.old_log_std ? expf(B.old_log_std[tid]) : sg;
  }
  __syncthreads();
  float logstd_sum = 0.0f;
#pragma unroll
  for (int k = 0; k < 12; ++k) logstd_sum += B.log_std[k];
  const float inv_b = 1.0f / (float)B.batch;
  float red[NRED];
#pragma unroll
  for (int i = 0; i < NRED; ++i) red[i] = 0.0f;

  for (long long i = (long long)blockIdx.x * THREADS + tid; i < B.batch; i += (long long)gridDim.x * THREADS) {
    float mu[12], act[12];
    ld12(B.mu + i * 12, mu);
    ld12(B.actions + i * 12, act);
    const float v = B.value[i], old_nlp = B.old_neglogp[i], adv = B.advantages[i], old_v = B.old_values[i], ret = B.returns[i];
    // ---- neglogp of the stored action under the fresh policy
    float u[12], ss = 0.0f;
#pragma unroll
    for (int k = 0; k < 12; ++k) { u[k] = (act[k] - mu[k]) / s_sigma[k]; ss += u[k] * u[k]; }
    const float nlp = 0.5f * ss + 11.027262398456072f + logstd_sum;
    // ---- actor loss: max(-adv ratio, -adv clamp(ratio, 1 - e, 1 + e))
    const float ratio = expf(old_nlp - nlp);
    const float clamped = fminf(fmaxf(ratio, 1.0f - B.e_clip), 1.0f + B.e_clip);
    const float s1 = -adv * ratio, s2 = -adv * clamped;
    const float a_loss = fmaxf(s1, s2);
    const bool inside = ratio >= 1.0f - B.e_clip && ratio <= 1.0f + B.e_clip;
    // d a_loss / d ratio: through surr1 when it is the larger (ties split evenly, and inside the range both carry -adv)
    const float dl_dratio = inside ? -adv : (s1 > s2 ? -adv : (s1 == s2 ? -0.5f * adv : 0.0f));
    const float dl_dnlp = dl_dratio * (-ratio);
    // ---- critic loss
    float c_loss, dc_dv;
    if (B.clip_value) {
      const float dv = v - old_v;
      const float vpc = old_v + fminf(fmaxf(dv, -B.e_clip), B.e_clip);
      const float l1 = (v - ret) * (v - ret), l2 = (vpc - ret) * (vpc - ret);
      const float g1 = 2.0f * (v - ret), g2 = (dv >= -B.e_clip && dv <= B.e_clip) ? 2.0f * (vpc - ret) : 0.0f;
      c_loss = fmaxf(l1, l2);
      dc_dv = l1 > l2 ? g1 : (l2 > l1 ? g2 : 0.5f * (g1 + g2));
    } else {
      c_loss = (ret - v) * (ret - v);
      dc_dv = 2.0f * (v - ret);
    }
    // ---- bound loss on the mean, gradients
    float b_loss = 0.0f, gmu[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const float hi = fmaxf(mu[k] - 1.1f, 0.0f), lo = fminf(mu[k] + 1.1f, 0.0f);
      b_loss += lo * lo + hi * hi;
      // d nlp / d mu_k = -(a_k - mu_k) / sigma_k^2
      gmu[k] = inv_b * (dl_dnlp * (-(act[k] - mu[k]) * s_inv_var[k]) + B.bounds_loss_coef * 2.0f * (hi + lo));
      // d nlp / d logstd_k = 1 - u_k^2
      red[5 + k] += dl_dnlp * (1.0f - u[k] * u[k]);
    }
#pragma unroll
    for (int q = 0; q < 3; ++q)
      __stcs(reinterpret_cast<float4*>(a.grad_mu + i * 12) + q, make_float4(gmu[4 * q], gmu[4 * q + 1], gmu[4 * q + 2], gmu[4 * q + 3]));
    a.grad_value[i] = inv_b * 0.5f * B.critic_coef * dc_dv;
    red[0] += a_loss; red[1] += c_loss; red[2] += b_loss;
    red[4] += inside ? 0.0f : 1.0f;
    if (B.old_mu) {       // rl_games policy_kl(p0 = fresh, p1 = old)
      float omu[12], kl = 0.0f;
      ld12(B.old_mu + i * 12, omu);
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const float p0s = s_sigma[k], p1s = s_old_sigma[k], dm = omu[k] - mu[k];
        kl += logf(p1s / p0s + 1e-5f) + (p0s * p0s + dm * dm) / (2.0f * (p1s * p1s + 1e-5f)) - 0.5f;
      }
      red[3] += kl;
      if (B.update_old) {   // rl_games dataset.update_mu_sigma: the next mini-epoch measures its KL against this pass's policy
        float4* om = reinterpret_cast<float4*>(const_cast<float*>(B.old_mu) + i * 12);
#pragma unroll
        for (int q = 0; q < 3; ++q) om[q] = make_float4(mu[4 * q], mu[4 * q + 1], mu[4 * q + 2], mu[4 * q + 3]);
      }
    }
  }
  // ---- CTA reduction (doubles), last CTA finishes
  {
    const double d = cta_sum<NRED, THREADS>(red, sm);
    if (tid < NRED) a.partials[(size_t)blockIdx.x * NRED + tid] = d;
  }
  if (!last_cta(a.ticket)) return;
  {
    __shared__ double part[12][NRED];
    const double d = grid_sum<NRED, 12>(a.partials, part);
    if (tid < NRED) sm[0][tid] = d;
  }
  __syncthreads();
  if (tid < 12) {
    // loss terms that reach log_std: actor (through neglogp) and -entropy_coef * entropy (d entropy / d logstd = 1)
    a.grad_log_std[tid] = (float)(sm[0][5 + tid] / (double)B.batch) - B.entropy_coef;
    // every CTA read old_log_std before it took its ticket: safe to overwrite in the last one
    if (B.update_old && B.old_log_std) const_cast<float*>(B.old_log_std)[tid] = B.log_std[tid];
  }
  if (tid == 0) {
    const double nb = (double)B.batch;
    const float a_l = (float)(sm[0][0] / nb), c_l = (float)(sm[0][1] / nb), b_l = (float)(sm[0][2] / nb);
    const float ent = 12.0f * (0.5f + 0.9189385332046727f) + logstd_sum;
    a.stats[0] = a_l; a.stats[1] = c_l; a.stats[2] = b_l; a.stats[3] = ent;
    a.stats[4] = (float)(sm[0][3] / nb);
    a.stats[5] = a_l + 0.5f * c_l * B.critic_coef - ent * B.entropy_coef + b_l * B.bounds_loss_coef;
    a.stats[6] = (float)(sm[0][4] / nb);
    a.stats[7] = 0.0f;
  }
}

// clip_grad_norm_ + Adam over a flat arena: a single CTA (the policies have 58,637 / 10,178 parameters)
constexpr int ADAM_THREADS = 1024;
__global__ void __launch_bounds__(ADAM_THREADS) k_adam_step(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                            int n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                                                            float max_norm, float grad_scale, float* __restrict__ norm_out, const LmAdamSchedule sched) {
  __shared__ double red[ADAM_THREADS / 32];
  __shared__ float s_coef, s_lr;
  const int tid = threadIdx.x;
  double ss = 0.0;
  for (int i0 = tid; i0 < n; i0 += 16 * ADAM_THREADS) {         // 16 independent loads in flight per thread (one CTA: latency-bound)
    float gv[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) { const int i = i0 + u * ADAM_THREADS; gv[u] = i < n ? g[i] * grad_scale : 0.0f; }
#pragma unroll
    for (int u = 0; u < 16; ++u) ss += (double)gv[u] * (double)gv[u];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < ADAM_THREADS / 32; ++w) t += red[w];
    const float total_norm = (float)sqrt(t);
    if (norm_out) *norm_out = total_norm;
    // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to <= 1
    s_coef = max_norm > 0.0f ? fminf(max_norm / (total_norm + 1e-6f), 1.0f) : 1.0f;
    // learning rate on the device (no host round trip per minibatch): this step runs with the stored rate, then rl_games'
    // AdaptiveScheduler.update moves it for the next one from the (rank-averaged) KL of this minibatch
    float cur = lr;
    if (sched.lr) {
      cur = *sched.lr;
      if (sched.kl) {
        const float kl = *sched.kl * sched.kl_scale;
        float nxt = cur;
        if (kl > 2.0f * sched.kl_threshold) nxt = fmaxf(cur / 1.5f, sched.lr_min);
        if (kl < 0.5f * sched.kl_threshold) nxt = fminf(cur * 1.5f, sched.lr_max);
        *sched.lr = nxt;
      }
    }
    s_lr = cur;
  }
  __syncthreads();
  const float coef = s_coef * grad_scale;
  const float step_size = s_lr / bc1;
  for (int i0 = tid; i0 < n; i0 += 8 * ADAM_THREADS) {          // 32 independent loads in flight per thread
    float gv[8], pv[8], mv[8], vv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * ADAM_THREADS;
      if (i < n) { gv[u] = g[i]; pv[u] = p[i]; mv[u] = m[i]; vv[u] = v[i]; }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * ADAM_THREADS;
      if (i >= n) continue;
      float gi = gv[u] * coef;
      g[i] = gi;                                                 // the clipped gradient, as clip_grad_norm_ leaves it
      if (wd != 0.0f) gi = fmaf(wd, pv[u], gi);                // torch.optim.Adam: grad = grad + weight_decay * param (not written back)
      const float mi = mv[u] + (gi - mv[u]) * (1.0f - beta1);  // exp_avg.lerp_(grad, 1 - beta1)
      const float vi = beta2 * vv[u] + (1.0f - beta2) * gi * gi;
      m[i] = mi; v[i] = vi;
      const float denom = sqrtf(vi) / bc2_sqrt + eps;
      p[i] = pv[u] - step_size * (mi / denom);
    }
  }
}

}  // namespace ppo
}  // namespace lm

using namespace lm;

void lm_rollout_workspace(LmRollout* r, double** partials, unsigned int** ticket, int* sm_count);   // lm_rollout.cu (library-internal)

extern "C" {

int lm_ppo_loss(LmRollout* workspace, const LmPpoBatch* batch, float* grad_mu, float* grad_value, float* grad_log_std, float* stats,
                void* stream) {
  if (!workspace || !batch || !grad_mu || !grad_value || !grad_log_std || !stats) return lm_fail(LM_ERR_BAD_ARG, "lm_ppo_loss: NULL argument");
  const LmPpoBatch& b = *batch;
  if (!b.mu || !b.value || !b.log_std || !b.actions || !b.old_neglogp || !b.advantages || !b.old_values || !b.returns || b.batch <= 0)
    return lm_fail(LM_ERR_BAD_ARG, "lm_ppo_loss: NULL batch tensor or empty batch");
  if ((b.old_mu == nullptr) != (b.old_log_std == nullptr)) return lm_fail(LM_ERR_BAD_ARG, "lm_ppo_loss: give old_mu and old_log_std or neither");
  if ((((uintptr_t)b.mu) | ((uintptr_t)b.actions) | ((uintptr_t)b.old_mu) | ((uintptr_t)grad_mu)) & 15)
    return lm_fail(LM_ERR_ALIGNMENT, "lm_ppo_loss: mu / actions / old_mu / grad_mu must be 16-byte aligned");
  ppo::Args a;
  a.b = b; a.grad_mu = grad_mu; a.grad_value = grad_value; a.grad_log_std = grad_log_std; a.stats = stats;
  int sm_count = 0;
  lm_rollout_workspace(workspace, &a.partials, &a.ticket, &sm_count);
  const long long want = (b.batch + ppo::THREADS - 1) / ppo::THREADS;
  const long long cap = (long long)sm_count * 4;                       // partials hold sm_count * 128 doubles >= 4 sm_count * NRED
  const int grid = (int)(want < cap ? want : cap);
  ppo::k_ppo_loss<<<grid, ppo::THREADS, 0, (cudaStream_t)stream>>>(a);
  return lm_check_launch("lm_ppo_loss");
}

int lm_adam_step_scheduled(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const LmAdamSchedule* sched, float lr,
                           float beta1, float beta2, float eps, float weight_decay, int32_t step, float max_grad_norm, float grad_scale,
                           float* grad_norm_out, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || step < 1) return lm_fail(LM_ERR_BAD_ARG, "lm_adam_step: bad argument");
  if (n > (1 << 24)) return lm_fail(LM_ERR_UNSUPPORTED, "lm_adam_step: single-CTA kernel, n <= 2^24");
  LmAdamSchedule sc;
  if (sched) sc = *sched; else { sc.lr = nullptr; sc.kl = nullptr; sc.kl_scale = 1.0f; sc.kl_threshold = 0.0f; sc.lr_min = 0.0f; sc.lr_max = 0.0f; }
  if (sc.kl && !sc.lr) return lm_fail(LM_ERR_BAD_ARG, "lm_adam_step_scheduled: a KL input needs the device learning rate");
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  ppo::k_adam_step<<<1, ppo::ADAM_THREADS, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, (int)n, lr, beta1, beta2, eps, weight_decay,
                                                                      bc1, bc2_sqrt, max_grad_norm, grad_scale, grad_norm_out, sc);
  return lm_check_launch("lm_adam_step");
}

int lm_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int32_t step, float max_grad_norm, float grad_scale, float* grad_norm_out, void* stream) {
  return lm_adam_step_scheduled(params, grads, exp_avg, exp_avg_sq, n, nullptr, lr, beta1, beta2, eps, weight_decay, step, max_grad_norm, grad_scale,
                                grad_norm_out, stream);
}

}  // extern "C"
