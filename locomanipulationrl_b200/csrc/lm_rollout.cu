// Rollout writer + running mean/std: the caller-side formats around the task/policy path.
//
// Not in the reference repo.  rl_games 1.6.0 (pinned at RobotLearning/setup.py:20; cfg/train/*PPO.yaml:46-47 normalize_input /
// normalize_value, :66 horizon) keeps an ExperienceBuffer of time-major tensors (obses, actions, mus, values, rewards, dones,
// neglogpacs) and RunningMeanStd modules; skrl (scripts/skrl_ppo_locomanipulation.py:100-122) a RandomMemory and
// RunningStandardScaler.  Both fold a batch into float64 running statistics with the parallel-variance rule
//     delta = mean_b - mean; tot = count + n; mean += delta n / tot; var = (var count + var_b n + delta^2 count n / tot) / tot
// where var_b is torch.var's unbiased batch variance.  Restated in oracle/port_policy.py.
//
// k_rollout_write: ONE launch per env step, one 1024-thread CTA per SM.  Phase A streams the observation tile [N,64] into obses[t]
// with 128-bit accesses; thread tid always owns channel quad tid % 16, so the per-channel moments (shifted by the current
// running mean against cancellation) accumulate in registers, meet per CTA in shared memory as doubles and per grid through
// the last-CTA pattern, where 64 threads apply the fold.  Phase B is per env: sample the action, its negative log-prob,
// and the scalar columns.  Algorithmic bytes: 368 read + 368 written per env-step (+48 for actions_out).
#include <string.h>

#include "lm_common.cuh"
#include "lm_host.h"
#include "lm_reduce.cuh"
#include "lm_rng.cuh"
#include "lm_sink.cuh"

namespace lm {
namespace rollout {

constexpr int THREADS = 1024;
constexpr int MAX_DIM = 64;
constexpr int PART = 2 * MAX_DIM;     // per-CTA partials: sum[64], sumsq[64]

struct Stats {
  double* mean; double* var; double* count;
  int dim; int variant;
};

struct WriteArgs {
  const float* obs; const float* mu; const float* value; const float* log_std; const float* noise; const float* rew;
  const long long* dones;
  float* obses; float* actions; float* mus; float* values; float* dones_f; float* neglogp; float* actions_out;
  // The reward that arrives with observation t is the reward of the action stored in slot t - 1 (rl_games play_steps: rewards[n] is
  // written after env_step(actions[n])): `rewards` / `prev_values` point at slot t - 1, NULL at t = 0.
  float* rewards; const float* prev_values; const long long* time_outs; float gamma;
  const float* value_denorm;            // [2] {mean, sqrt(var + eps)} of the value normaliser, NULL: values stored as they come
  LmActionSink sink; int has_sink;      // the sampled actions also go straight into the task's buffers (lm_sink.cuh)
  int n; float reward_scale;
  unsigned long long noise_seed, noise_step; int sample_in_kernel;
  Stats st; int has_stats;
  double* partials; unsigned int* ticket;
};

// fold of a whole batch (n samples, shifted sums S1 = sum(x - shift), S2 = sum((x - shift)^2)) into the running statistics
__device__ __forceinline__ void fold(const Stats& st, int c, double shift, double S1, double S2, double n) {
  const double mean_b = shift + S1 / n;
  const double var_b = n > 1.0 ? (S2 - S1 * S1 / n) / (n - 1.0) : 0.0;      // torch.var: unbiased
  const double mean = st.mean[c], var = st.var[c], count = st.count[0];
  const double delta = mean_b - mean, tot = count + n;
  st.mean[c] = mean + delta * n / tot;
  st.var[c] = (var * count + var_b * n + delta * delta * count * n / tot) / tot;
}

// block-level meeting point of the per-thread moments: thread tid owns channels 4 (tid % 16) .. +3
__device__ __forceinline__ void cta_moments(const float s1[4], const float s2[4], double* sm /* [THREADS/32 warps][16 quads][8] */, double* out /* [PART] */) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double v[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) { v[k] = (double)s1[k]; v[4 + k] = (double)s2[k]; }
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);      // lanes l and l+16 share a channel quad
  if (lane < 16) {
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[(warp * 16 + lane) * 8 + k] = v[k];
  }
  __syncthreads();
  if (tid < PART) {                    // tid = which * 64 + channel
    const int which = tid >> 6, c = tid & 63, quad = c >> 2, k = (c & 3) + 4 * which;
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) s += sm[(w * 16 + quad) * 8 + k];
    out[tid] = s;
  }
}

// grid-wide sum of partials[cta][PART] -> sm_tot[PART] (deterministic order), by the last CTA
__device__ __forceinline__ void grid_moments(const double* partials, double* sm /* >= 4 * PART */, double* sm_tot) {
  constexpr int SUB = THREADS / PART;                                      // 8 subsets of CTAs x 128 values
  constexpr int MAXC = 24;                                                 // loads per thread, all in flight: grids up to 192 CTAs
  const int tid = threadIdx.x, sub = tid / PART, j = tid & (PART - 1);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int c0 = sub; c0 < (int)gridDim.x; c0 += SUB * MAXC) {
    double v[MAXC];
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
      const int cta = c0 + SUB * u;
      v[u] = cta < (int)gridDim.x ? __ldcg(partials + (size_t)cta * PART + j) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < MAXC; ++u) acc[u & 3] += v[u];
  }
  sm[sub * PART + j] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  __syncthreads();
  if (tid < PART) {
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < SUB; ++u) s += sm[u * PART + tid];
    sm_tot[tid] = s;
  }
  __syncthreads();
}

// rl_games RunningMeanStd(denorm=True): sqrt(var + eps) * clamp(v, -5, 5) + mean (the network's value head is trained in
// normalised units when normalize_value is on; the rollout stores de-normalised values)
__device__ __forceinline__ float denorm_value(const float* dn, float v) {
  return dn ? add(mul(dn[1], fminf(fmaxf(v, -5.0f), 5.0f)), dn[0]) : v;
}
// rewards_shaper scale, then value_bootstrap: + gamma * value * time_out (a2c_common.play_steps)
template <typename A>
__device__ __forceinline__ float shaped_reward(const A& a, unsigned int e, float rw) {
  float r = mul(rw, a.reward_scale);
  if (a.time_outs && a.time_outs[e] != 0) r = add(r, mul(a.gamma, a.prev_values[e]));
  return r;
}
struct EnvIn { float4 mv[3], nv[3]; float val, rw; long long dn; };
__device__ __forceinline__ EnvIn env_load(const WriteArgs& a, unsigned int e) {
  EnvIn in;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    in.mv[q] = __ldg(reinterpret_cast<const float4*>(a.mu + e * 12) + q);            // rows are 48 B: let L1 keep the shared sectors
    in.nv[q] = a.noise ? __ldg(reinterpret_cast<const float4*>(a.noise + e * 12) + q)
                       : (a.sample_in_kernel ? normal4(philox_at(a.noise_seed, a.noise_step, (unsigned long long)e * 3ull + q)) : make_float4(0.f, 0.f, 0.f, 0.f));
  }
  in.val = a.value[e]; in.rw = a.rewards ? a.rew[e] : 0.0f; in.dn = a.dones[e];
  return in;
}
__device__ __forceinline__ void env_emit(const WriteArgs& a, unsigned int e, const EnvIn& in, const float* s_sigma, float lsum) {
  float ss = 0.0f;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const float4 sg = *reinterpret_cast<const float4*>(&s_sigma[4 * q]);
    const float m[4] = {in.mv[q].x, in.mv[q].y, in.mv[q].z, in.mv[q].w}, z[4] = {in.nv[q].x, in.nv[q].y, in.nv[q].z, in.nv[q].w};
    const float sgm[4] = {sg.x, sg.y, sg.z, sg.w};
    float act[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      act[k] = add(mul(z[k], sgm[k]), m[k]);                    // torch.normal: randn * std + mean
      const float u = fdiv(sub(act[k], m[k]), sgm[k]);          // the libraries divide: ((x - mean) / std)^2
      ss += u * u;
    }
    const float4 av = make_float4(act[0], act[1], act[2], act[3]);
    __stcs(reinterpret_cast<float4*>(a.actions + e * 12) + q, av);
    __stcs(reinterpret_cast<float4*>(a.mus + e * 12) + q, in.mv[q]);
    if (a.actions_out) reinterpret_cast<float4*>(a.actions_out + e * 12)[q] = av;
    if (a.has_sink) sink_store(a.sink, (int)e, q, av);
  }
  a.neglogp[e] = 0.5f * ss + 11.027262398456072f + lsum;      // 0.5 log(2 pi) * 12
  a.values[e] = denorm_value(a.value_denorm, in.val);
  if (a.rewards) a.rewards[e] = shaped_reward(a, e, in.rw);
  a.dones_f[e] = in.dn != 0 ? 1.0f : 0.0f;
}

__global__ void __launch_bounds__(THREADS, 1) k_rollout_write(const WriteArgs a) {
  __shared__ double sm[(THREADS / 32) * 16 * 8];
  __shared__ double sm_tot[PART];
  const int tid = threadIdx.x;
  const unsigned int gtid = blockIdx.x * THREADS + tid, gstride = gridDim.x * THREADS;     // n <= 2^27 (host check): 32-bit indices
  __shared__ __align__(16) float s_sigma[12];
  __shared__ float s_lsum;
  pdl_launch_dependents();
  pdl_wait();                                                          // log_std may be the predecessor's output (optimiser step)
  if (tid < 12) s_sigma[tid] = (float)exp((double)a.log_std[tid]);     // correctly rounded; torch's exp is within 1 ulp of it
  if (tid == 32) {
    float ls = 0.0f;
    for (int k = 0; k < 12; ++k) ls += a.log_std[k];
    s_lsum = ls;
  }
  __syncthreads();
  // ---- phase A: observations -> obses[t]; per-channel shifted moments
  {
    const int c4 = tid & 15;
    float shift[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.has_stats) {
#pragma unroll
      for (int k = 0; k < 4; ++k) shift[k] = (float)a.st.mean[4 * c4 + k];
    }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    const unsigned int total = (unsigned int)a.n * 16u;
    const float4* src = reinterpret_cast<const float4*>(a.obs);
    float4* dst = reinterpret_cast<float4*>(a.obses);
    constexpr int SLOTS = 8;                     // 65,536 envs on 148 x 1024 threads: 6.9 float4 per thread, all in flight at once
    for (unsigned int base = gtid; base < total; base += SLOTS * gstride) {
      float4 v[SLOTS];
#pragma unroll
      for (int u = 0; u < SLOTS; ++u) {
        const unsigned int i = base + u * gstride;
        if (i < total) v[u] = ld_nc_f4(reinterpret_cast<const float*>(src + i));
      }
#pragma unroll
      for (int u = 0; u < SLOTS; ++u) {
        const unsigned int i = base + u * gstride;
        if (i < total) {
          __stcs(dst + i, v[u]);
          const float d[4] = {v[u].x - shift[0], v[u].y - shift[1], v[u].z - shift[2], v[u].w - shift[3]};
#pragma unroll
          for (int k = 0; k < 4; ++k) { s1[k] += d[k]; s2[k] = fmaf(d[k], d[k], s2[k]); }
        }
      }
    }
    // phase B's loads for this thread's first env are issued before the CTA-level meeting point of the moments
    EnvIn in0;
    const bool have0 = gtid < (unsigned int)a.n;
    if (have0) in0 = env_load(a, gtid);
    if (a.has_stats) cta_moments(s1, s2, sm, a.partials + (size_t)blockIdx.x * PART);
    // ---- phase B: per env: action sample, negative log-prob, scalar columns
    const float lsum = s_lsum;
    if (have0) env_emit(a, gtid, in0, s_sigma, lsum);
    for (unsigned int e = gtid + gstride; e < (unsigned int)a.n; e += gstride) env_emit(a, e, env_load(a, e), s_sigma, lsum);
  }
  // ---- fold the batch into the running statistics (last CTA)
  if (a.has_stats && last_cta(a.ticket)) {
    grid_moments(a.partials, sm, sm_tot);
    if (tid < a.st.dim) fold(a.st, tid, (double)(float)a.st.mean[tid], sm_tot[tid], sm_tot[MAX_DIM + tid], (double)a.n);
    __syncthreads();
    if (tid == 0) a.st.count[0] += (double)a.n;
  }
}

// closes a rollout: the reward of the last stored action (slot T - 1) and the bootstrap inputs of GAE (last value, last done)
struct FinishArgs {
  const float* rew; const float* value; const long long* dones; const long long* time_outs;
  float* rewards; const float* prev_values; float* last_values; float* last_dones;
  const float* value_denorm; float gamma; float reward_scale; int n;
};
__global__ void k_rollout_finish(const FinishArgs a) {
  const unsigned int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (unsigned int)a.n) return;
  a.rewards[e] = shaped_reward(a, e, a.rew[e]);
  a.last_values[e] = denorm_value(a.value_denorm, a.value[e]);
  a.last_dones[e] = a.dones[e] != 0 ? 1.0f : 0.0f;
}

// y = clamp((x - mean) * inv_std, -clip, clip)   (RunningMeanStd.forward in eval mode; clip <= 0: no clamp)
__global__ void k_normalize_clamp(const float* __restrict__ x, float* __restrict__ y, long long n, const float* __restrict__ mean,
                                  const float* __restrict__ inv_std, float clip) {
  const float m = mean[0], is = inv_std[0];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = mul(sub(x[i], m), is);
    if (clip > 0.0f) v = fminf(fmaxf(v, -clip), clip);
    y[i] = v;
  }
}

// moments[0..2] = {sum(x), sum(x^2), n} as doubles (advantage normalisation; summed over ranks by the caller when sharded)
__global__ void __launch_bounds__(THREADS, 1) k_moments(const float* __restrict__ x, long long n, double* __restrict__ moments, double* partials,
                                                        unsigned int* ticket) {
  __shared__ double sm[THREADS / 32][2];
  const int tid = threadIdx.x;
  const long long stride = (long long)gridDim.x * THREADS;
  double s1 = 0.0, s2 = 0.0;
  for (long long i = (long long)blockIdx.x * THREADS + tid; i < n; i += stride) { const double v = (double)x[i]; s1 += v; s2 += v * v; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, off); s2 += __shfl_xor_sync(0xffffffffu, s2, off); }
  if ((tid & 31) == 0) { sm[tid >> 5][0] = s1; sm[tid >> 5][1] = s2; }
  __syncthreads();
  if (tid < 2) {
    double s = 0.0;
    for (int w = 0; w < THREADS / 32; ++w) s += sm[w][tid];
    partials[(size_t)blockIdx.x * 2 + tid] = s;
  }
  if (last_cta(ticket)) {
    if (tid < 2) {
      double s = 0.0;
      for (int c = 0; c < (int)gridDim.x; ++c) s += __ldcg(partials + (size_t)c * 2 + tid);
      moments[tid] = s;
    }
    if (tid == 2) moments[2] = (double)n;
  }
}
// y = (x - mean) / (std + eps), mean / unbiased std from the (possibly all-reduced) moments: torch `(a - a.mean()) / (a.std() + 1e-8)`
__global__ void k_moment_normalize(const float* __restrict__ x, float* __restrict__ y, long long n, const double* __restrict__ moments, float eps) {
  const double cnt = moments[2], mean_d = moments[0] / cnt;
  const double var = cnt > 1.0 ? fmax((moments[1] - moments[0] * mean_d) / (cnt - 1.0), 0.0) : 0.0;
  const float mean = (float)mean_d, denom = add((float)sqrt(var), eps);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = fdiv(sub(x[i], mean), denom);
}

// stand-alone fold of x [n, dim] (dim = 1 or a multiple of 4, <= 64)
__global__ void __launch_bounds__(THREADS, 1) k_rms_update(const float* __restrict__ x, long long n, Stats st, double* partials, unsigned int* ticket) {
  __shared__ double sm[(THREADS / 32) * 16 * 8];
  __shared__ double sm_tot[PART];
  const int tid = threadIdx.x;
  const size_t gtid = (size_t)blockIdx.x * THREADS + tid, gstride = (size_t)gridDim.x * THREADS;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (st.dim == 1) {
    // scalar column: 128-bit loads, four of them in flight per thread; every thread sums into its own pair, merged below
    const float shift = (float)st.mean[0];
    const size_t n4 = (((uintptr_t)x) & 15) == 0 ? (size_t)n >> 2 : 0;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    size_t i = gtid;
    for (; i + 3 * gstride < n4; i += 4 * gstride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld_nc_f4(reinterpret_cast<const float*>(x4 + i + u * gstride));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float d[4] = {v[u].x - shift, v[u].y - shift, v[u].z - shift, v[u].w - shift};
#pragma unroll
        for (int k = 0; k < 4; ++k) { s1[0] += d[k]; s2[0] = fmaf(d[k], d[k], s2[0]); }
      }
    }
    for (; i < n4; i += gstride) {
      const float4 v = ld_nc_f4(reinterpret_cast<const float*>(x4 + i));
      const float d[4] = {v.x - shift, v.y - shift, v.z - shift, v.w - shift};
#pragma unroll
      for (int k = 0; k < 4; ++k) { s1[0] += d[k]; s2[0] = fmaf(d[k], d[k], s2[0]); }
    }
    for (size_t j = (n4 << 2) + gtid; j < (size_t)n; j += gstride) {
      const float d = x[j] - shift;
      s1[0] += d; s2[0] = fmaf(d, d, s2[0]);
    }
  } else {
    // dim = 4 q: float4 index i covers channels 4 (i % q) ..; THREADS and the grid stride are multiples of 16 only, so the quad of a
    // thread changes from iteration to iteration unless q divides the stride: walk rows instead (thread -> fixed quad, rows strided)
    const int q = st.dim >> 2;
    const int quad = (int)(gtid % q);
    float shift[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) shift[k] = (float)st.mean[4 * quad + k];
    const size_t rows_per_pass = gstride / q;       // threads [0, rows_per_pass * q) are active
    if (gtid < rows_per_pass * q) {
      // four rows (128-bit loads) in flight per thread, accumulated in row order (the sums are those of the one-load loop, bit for
      // bit; with one load in flight the kernel ran at 0.27 of the HBM rate: 57.7 us per 393,216 x 64 minibatch)
      size_t row = gtid / q;
      for (; row + 3 * rows_per_pass < (size_t)n; row += 4 * rows_per_pass) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_nc_f4(x + ((row + u * rows_per_pass) * q + quad) * 4);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float d[4] = {v[u].x - shift[0], v[u].y - shift[1], v[u].z - shift[2], v[u].w - shift[3]};
#pragma unroll
          for (int k = 0; k < 4; ++k) { s1[k] += d[k]; s2[k] = fmaf(d[k], d[k], s2[k]); }
        }
      }
      for (; row < (size_t)n; row += rows_per_pass) {
        const float4 v = ld_nc_f4(x + (row * q + quad) * 4);
        const float d[4] = {v.x - shift[0], v.y - shift[1], v.z - shift[2], v.w - shift[3]};
#pragma unroll
        for (int k = 0; k < 4; ++k) { s1[k] += d[k]; s2[k] = fmaf(d[k], d[k], s2[k]); }
      }
    }
    // move the sums to the slot of the thread's quad: cta_moments() keys on tid % 16, so publish through shared memory instead
    __shared__ double qsum[16][8];
    if (tid < 16 * 8) qsum[tid >> 3][tid & 7] = 0.0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(&qsum[quad][k], (double)s1[k]);       // shared-memory fp64 atomics: order-dependent only below 1e-16 relative
      atomicAdd(&qsum[quad][4 + k], (double)s2[k]);
    }
    __syncthreads();
    if (tid < PART) {
      const int which = tid >> 6, c = tid & 63;
      partials[(size_t)blockIdx.x * PART + tid] = (c >> 2) < q ? qsum[c >> 2][(c & 3) + 4 * which] : 0.0;
    }
  }
  if (st.dim == 1) {
    double v1 = (double)s1[0], v2 = (double)s2[0];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { v1 += __shfl_xor_sync(0xffffffffu, v1, off); v2 += __shfl_xor_sync(0xffffffffu, v2, off); }
    if ((tid & 31) == 0) { sm[(tid >> 5) * 2] = v1; sm[(tid >> 5) * 2 + 1] = v2; }
    __syncthreads();
    if (tid < PART) {
      double s = 0.0;
      if (tid == 0 || tid == MAX_DIM) {
        for (int w = 0; w < THREADS / 32; ++w) s += sm[w * 2 + (tid == 0 ? 0 : 1)];
      }
      partials[(size_t)blockIdx.x * PART + tid] = s;
    }
  }
  if (last_cta(ticket)) {
    grid_moments(partials, sm, sm_tot);
    if (tid < st.dim) fold(st, tid, (double)(float)st.mean[tid], sm_tot[tid], sm_tot[MAX_DIM + tid], (double)n);
    __syncthreads();
    if (tid == 0) st.count[0] += (double)n;
  }
}

__global__ void k_rms_export(Stats st, float* __restrict__ mean_f32, float* __restrict__ inv_std_f32) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= st.dim) return;
  mean_f32[c] = (float)st.mean[c];
  const float var = (float)st.var[c];
  // rl_games RunningMeanStd: / sqrt(var.float() + 1e-5);  skrl RunningStandardScaler: / (sqrt(var.float()) + 1e-8)
  inv_std_f32[c] = st.variant == LM_RMS_RLGAMES ? fdiv(1.0f, __fsqrt_rn(add(var, 1e-5f))) : fdiv(1.0f, add(__fsqrt_rn(var), 1e-8f));
}

// {mean, std} of a scalar normaliser in the form RunningMeanStd(denorm=True) applies: y = std * clamp(x, -5, 5) + mean
__global__ void k_rms_export_denorm(Stats st, float* __restrict__ out) {
  if (threadIdx.x != 0) return;
  out[0] = (float)st.mean[0];
  const float var = (float)st.var[0];
  out[1] = st.variant == LM_RMS_RLGAMES ? __fsqrt_rn(add(var, 1e-5f)) : add(__fsqrt_rn(var), 1e-8f);
}

}  // namespace rollout
}  // namespace lm

using namespace lm;
using namespace lm::rollout;

struct LmRollout {
  int device;
  int sm_count;
  double* partials;        // [sm_count][PART]
  unsigned int* ticket;
};

// workspace shared with lm_ppo.cu (partials: sm_count * 128 doubles; ticket: zero between launches)
void lm_rollout_workspace(LmRollout* r, double** partials, unsigned int** ticket, int* sm_count) {
  *partials = r->partials; *ticket = r->ticket; *sm_count = r->sm_count;
}

static int check_stats(const LmRunningStats* s, const char* who) {
  if (!s->mean || !s->var || !s->count) return lm_fail(LM_ERR_BAD_ARG, "%s: NULL running statistics", who);
  if (!(s->dim == 1 || (s->dim % 4 == 0 && s->dim >= 4 && s->dim <= MAX_DIM))) return lm_fail(LM_ERR_BAD_ARG, "%s: dim must be 1 or a multiple of 4 <= 64", who);
  if (s->variant != LM_RMS_RLGAMES && s->variant != LM_RMS_SKRL) return lm_fail(LM_ERR_BAD_ARG, "%s: unknown variant", who);
  return LM_OK;
}

extern "C" {

LmRollout* lm_rollout_create(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { lm_set_error("lm_rollout_create: %s", cudaGetErrorString(e)); return nullptr; }
  if (prop.major != 10) { lm_set_error("lm_rollout_create: library is sm_100a only (device sm_%d%d)", prop.major, prop.minor); return nullptr; }
  cudaSetDevice(device);
  LmRollout* r = new LmRollout();
  r->device = device; r->sm_count = prop.multiProcessorCount; r->partials = nullptr; r->ticket = nullptr;
  if (cudaMalloc(&r->partials, (size_t)r->sm_count * PART * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&r->ticket, sizeof(unsigned int)) != cudaSuccess) {
    lm_set_error("lm_rollout_create: cudaMalloc failed");
    cudaFree(r->partials); delete r; return nullptr;
  }
  cudaMemset(r->ticket, 0, sizeof(unsigned int));
  return r;
}

void lm_rollout_destroy(LmRollout* r) {
  if (!r) return;
  cudaFree(r->partials); cudaFree(r->ticket);
  delete r;
}

int lm_rollout_write(LmRollout* r, const LmRolloutStep* step, const LmRolloutBuffers* buf, int32_t t, float* actions_out,
                     const LmRunningStats* obs_stats, void* stream) {
  if (!r || !step || !buf) return lm_fail(LM_ERR_BAD_ARG, "lm_rollout_write: NULL argument");
  if (!step->obs || !step->mu || !step->value || !step->log_std || !step->dones) return lm_fail(LM_ERR_BAD_ARG, "lm_rollout_write: NULL step tensor");
  if (t > 0 && !step->rew) return lm_fail(LM_ERR_BAD_ARG, "lm_rollout_write: rew is required for t > 0 (it completes slot t - 1)");
  if (!buf->obses || !buf->actions || !buf->mus || !buf->values || !buf->rewards || !buf->dones || !buf->neglogp) return lm_fail(LM_ERR_BAD_ARG, "lm_rollout_write: NULL rollout buffer");
  if (buf->num_envs > (1 << 27)) return lm_fail(LM_ERR_UNSUPPORTED, "lm_rollout_write: more than 2^27 envs per call");
  if (buf->num_envs <= 0 || t < 0 || t >= buf->horizon) return lm_fail(LM_ERR_BAD_ARG, "lm_rollout_write: t = %d outside [0, %d)", t, buf->horizon);
  if ((((uintptr_t)step->obs) | ((uintptr_t)step->mu) | ((uintptr_t)step->noise) | ((uintptr_t)buf->obses) | ((uintptr_t)buf->actions) | ((uintptr_t)buf->mus) |
       ((uintptr_t)actions_out)) & 15)
    return lm_fail(LM_ERR_ALIGNMENT, "lm_rollout_write: obs / mu / noise / obses / actions / mus must be 16-byte aligned");
  if (obs_stats) {
    const int rc = check_stats(obs_stats, "lm_rollout_write");
    if (rc != LM_OK) return rc;
    if (obs_stats->dim != 64) return lm_fail(LM_ERR_BAD_ARG, "lm_rollout_write: observation statistics must have dim 64");
  }
  const size_t n = (size_t)buf->num_envs;
  WriteArgs a;
  a.obs = step->obs; a.mu = step->mu; a.value = step->value; a.log_std = step->log_std; a.noise = step->noise; a.rew = step->rew;
  a.dones = reinterpret_cast<const long long*>(step->dones);
  a.obses = buf->obses + (size_t)t * n * 64; a.actions = buf->actions + (size_t)t * n * 12; a.mus = buf->mus + (size_t)t * n * 12;
  a.values = buf->values + (size_t)t * n; a.dones_f = buf->dones + (size_t)t * n;
  a.rewards = t > 0 ? buf->rewards + (size_t)(t - 1) * n : nullptr;
  a.prev_values = t > 0 ? buf->values + (size_t)(t - 1) * n : nullptr;
  a.time_outs = t > 0 ? reinterpret_cast<const long long*>(step->time_outs) : nullptr; a.gamma = step->gamma;
  a.value_denorm = step->value_denorm;
  a.has_sink = step->sink != nullptr;
  if (step->sink) a.sink = *step->sink; else memset(&a.sink, 0, sizeof(a.sink));
  a.neglogp = buf->neglogp + (size_t)t * n; a.actions_out = actions_out;
  a.n = buf->num_envs; a.reward_scale = buf->reward_scale;
  a.noise_seed = step->noise_seed; a.noise_step = step->noise_step; a.sample_in_kernel = step->sample_in_kernel;
  a.has_stats = obs_stats ? 1 : 0;
  if (obs_stats) a.st = Stats{obs_stats->mean, obs_stats->var, obs_stats->count, obs_stats->dim, obs_stats->variant};
  else a.st = Stats{nullptr, nullptr, nullptr, 0, 0};
  a.partials = r->partials; a.ticket = r->ticket;
  const size_t work = (n * 16 + THREADS - 1) / THREADS;
  const int grid = (int)(work < (size_t)r->sm_count ? work : (size_t)r->sm_count);
  cudaError_t le = lm_launch(k_rollout_write, dim3(grid), dim3(THREADS), 0, (cudaStream_t)stream, true, a);
  if (le != cudaSuccess) return lm_fail((int)le, "lm_rollout_write: %s", cudaGetErrorString(le));
  return lm_check_launch("lm_rollout_write");
}

int lm_rollout_finish(LmRollout* r, const LmRolloutStep* step, const LmRolloutBuffers* buf, float* last_values, float* last_dones, void* stream) {
  if (!r || !step || !buf || !last_values || !last_dones) return lm_fail(LM_ERR_BAD_ARG, "lm_rollout_finish: NULL argument");
  if (!step->rew || !step->value || !step->dones || !buf->rewards || !buf->values || buf->num_envs <= 0 || buf->horizon <= 0)
    return lm_fail(LM_ERR_BAD_ARG, "lm_rollout_finish: NULL tensor");
  const size_t n = (size_t)buf->num_envs, last = (size_t)(buf->horizon - 1) * n;
  FinishArgs a;
  a.rew = step->rew; a.value = step->value; a.dones = reinterpret_cast<const long long*>(step->dones);
  a.time_outs = reinterpret_cast<const long long*>(step->time_outs);
  a.rewards = buf->rewards + last; a.prev_values = buf->values + last; a.last_values = last_values; a.last_dones = last_dones;
  a.value_denorm = step->value_denorm; a.gamma = step->gamma; a.reward_scale = buf->reward_scale; a.n = buf->num_envs;
  k_rollout_finish<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  return lm_check_launch("lm_rollout_finish");
}

int lm_normalize_clamp(const float* x, float* y, int64_t n, const float* mean, const float* inv_std, float clip, void* stream) {
  if (!x || !y || !mean || !inv_std || n <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_normalize_clamp: bad argument");
  const long long want = (n + 255) / 256;
  k_normalize_clamp<<<(unsigned)(want < 4096 ? want : 4096), 256, 0, (cudaStream_t)stream>>>(x, y, (long long)n, mean, inv_std, clip);
  return lm_check_launch("lm_normalize_clamp");
}

int lm_moments(LmRollout* r, const float* x, int64_t n, double* moments, void* stream) {
  if (!r || !x || !moments || n <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_moments: bad argument");
  const size_t work = ((size_t)n + THREADS - 1) / THREADS;
  const int grid = (int)(work < (size_t)r->sm_count ? work : (size_t)r->sm_count);
  k_moments<<<grid, THREADS, 0, (cudaStream_t)stream>>>(x, (long long)n, moments, r->partials, r->ticket);
  return lm_check_launch("lm_moments");
}

int lm_moment_normalize(const float* x, float* y, int64_t n, const double* moments, float eps, void* stream) {
  if (!x || !y || !moments || n <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_moment_normalize: bad argument");
  const long long want = (n + 255) / 256;
  k_moment_normalize<<<(unsigned)(want < 4096 ? want : 4096), 256, 0, (cudaStream_t)stream>>>(x, y, (long long)n, moments, eps);
  return lm_check_launch("lm_moment_normalize");
}

int lm_rms_update(LmRollout* r, const float* x, int64_t n, const LmRunningStats* stats, void* stream) {
  if (!r || !x || !stats || n <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_rms_update: bad argument");
  const int rc = check_stats(stats, "lm_rms_update");
  if (rc != LM_OK) return rc;
  if (stats->dim > 1 && (((uintptr_t)x) & 15)) return lm_fail(LM_ERR_ALIGNMENT, "lm_rms_update: x must be 16-byte aligned");
  const size_t items = stats->dim == 1 ? (size_t)n : (size_t)n * (stats->dim / 4);
  const size_t work = (items + THREADS - 1) / THREADS;
  const int grid = (int)(work < (size_t)r->sm_count ? work : (size_t)r->sm_count);
  Stats st{stats->mean, stats->var, stats->count, stats->dim, stats->variant};
  k_rms_update<<<grid, THREADS, 0, (cudaStream_t)stream>>>(x, (long long)n, st, r->partials, r->ticket);
  return lm_check_launch("lm_rms_update");
}

int lm_rms_export_denorm(const LmRunningStats* stats, float* mean_std_f32, void* stream) {
  if (!stats || !mean_std_f32) return lm_fail(LM_ERR_BAD_ARG, "lm_rms_export_denorm: NULL argument");
  const int rc = check_stats(stats, "lm_rms_export_denorm");
  if (rc != LM_OK) return rc;
  if (stats->dim != 1) return lm_fail(LM_ERR_BAD_ARG, "lm_rms_export_denorm: scalar statistics only (the value normaliser)");
  Stats st{stats->mean, stats->var, stats->count, stats->dim, stats->variant};
  k_rms_export_denorm<<<1, 32, 0, (cudaStream_t)stream>>>(st, mean_std_f32);
  return lm_check_launch("lm_rms_export_denorm");
}

int lm_rms_export(const LmRunningStats* stats, float* mean_f32, float* inv_std_f32, void* stream) {
  if (!stats || !mean_f32 || !inv_std_f32) return lm_fail(LM_ERR_BAD_ARG, "lm_rms_export: NULL argument");
  const int rc = check_stats(stats, "lm_rms_export");
  if (rc != LM_OK) return rc;
  Stats st{stats->mean, stats->var, stats->count, stats->dim, stats->variant};
  k_rms_export<<<1, 64, 0, (cudaStream_t)stream>>>(st, mean_f32, inv_std_f32);
  return lm_check_launch("lm_rms_export");
}

}  // extern "C"
