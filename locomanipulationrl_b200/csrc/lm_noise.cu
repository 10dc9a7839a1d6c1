// Observation / action noise of the vec-env step.
// Reference: utils/domain_randomization/randomize.py:212-305 (apply_observations_randomization, apply_actions_randomization,
// _apply_correlated_noise, _apply_uncorrelated_noise), called from envs/vec_env_rlgames.py:57-58,70-72.
//
// The reference compacts the affected envs with nonzero() (a host synchronisation per buffer per step) and draws (k, dim)
// tensors.  Here one launch handles a buffer: a tile of 256 envs computes its flags (and, when the caller supplies draws
// in the reference's compact order, the ranks: per-tile counts from a first small launch + a block scan), then the tile's
// [256, dim] span is walked with 128-bit accesses.  Without supplied draws the noise comes from a counter-based
// generator keyed by (seed, step, element), so the kernel needs neither draw tensors nor ranks.
#include "lm_common.cuh"
#include "lm_host.h"
#include "lm_rng.cuh"

namespace lm {
namespace noise {

constexpr int TILE = 256;       // envs (= threads) per CTA: 256 CTAs at 65,536 envs, several resident per SM

struct Args {
  LmNoiseState s;
  float* buffer;
  const long long* reset_buf;
  const float* draws_reset;
  const float* draws_interval;
  int compact;
  unsigned long long seed, step;
  int* counts;          // [tiles][2] (compact mode)
};

// counter / flags of env e exactly as randomize.py:214-216,229-230 (and :239-241,254-255) update them
__device__ __forceinline__ void env_flags(const LmNoiseState& s, const long long* reset_buf, int e, int& counter_out, int& flag_r, int& flag_i) {
  flag_r = reset_buf[e] != 0;
  int c = flag_r ? 0 : s.counter[e];
  c += 1;
  flag_i = s.on_interval.enabled && c >= s.on_interval.frequency_interval;
  if (flag_i) c = 0;
  counter_out = c;
}

__global__ void __launch_bounds__(TILE) k_noise_count(const Args a) {
  __shared__ int sm[2][TILE / 32];
  const int e = blockIdx.x * TILE + threadIdx.x;
  int c, fr = 0, fi = 0;
  if (e < a.s.num_envs) env_flags(a.s, a.reset_buf, e, c, fr, fi);
  const unsigned br = __ballot_sync(0xffffffffu, fr), bi = __ballot_sync(0xffffffffu, fi);
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = __popc(br); sm[1][threadIdx.x >> 5] = __popc(bi); }
  __syncthreads();
  if (threadIdx.x < 2) {
    int t = 0;
    for (int w = 0; w < TILE / 32; ++w) t += sm[threadIdx.x][w];
    a.counts[blockIdx.x * 2 + threadIdx.x] = t;
  }
}

__device__ __forceinline__ float4 draw4(const Args& a, int distribution, unsigned long long element, unsigned int which) {
  const uint4 r = philox_at(a.seed, a.step * 2ull + which, element);
  return distribution == LM_DIST_GAUSSIAN ? normal4(r) : uniform4(r);
}
__device__ __forceinline__ float shape1(const LmNoiseSpec& sp, float d) {
  const float v = add(mul(sp.scale, d), sp.shift);          // randn * std + mean | (hi - lo) * rand + lo | (log hi - log lo) * rand + log lo
  return sp.distribution == LM_DIST_LOGUNIFORM ? expf(v) : v;
}
__device__ __forceinline__ float4 shape4(const LmNoiseSpec& sp, const float4& d) {
  return make_float4(shape1(sp, d.x), shape1(sp, d.y), shape1(sp, d.z), shape1(sp, d.w));
}
__device__ __forceinline__ float4 op4(int operation, const float4& x, const float4& n) {
  return operation == LM_NOISE_ADDITIVE ? make_float4(add(x.x, n.x), add(x.y, n.y), add(x.z, n.z), add(x.w, n.w))
                                        : make_float4(mul(x.x, n.x), mul(x.y, n.y), mul(x.z, n.z), mul(x.w, n.w));
}

__global__ void __launch_bounds__(TILE) k_noise_apply(const Args a) {
  __shared__ int row_r[TILE], row_i[TILE];        // draw row of the env (-1: not affected)
  __shared__ int warp_tot[2][TILE / 32];
  __shared__ int base[2];
  const LmNoiseState& s = a.s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int e = blockIdx.x * TILE + tid;
  int c = 0, fr = 0, fi = 0;
  if (e < s.num_envs) {
    env_flags(s, a.reset_buf, e, c, fr, fi);
    s.counter[e] = c;
  }
  int rr = e, ri = e;                             // dense rows: env e reads row e
  if (a.compact) {
    // ranks among the affected envs in ascending order: preceding tiles (k_noise_count) + block scan
    int t0 = 0, t1 = 0;
    for (int b = tid; b < (int)blockIdx.x; b += TILE) { t0 += a.counts[b * 2]; t1 += a.counts[b * 2 + 1]; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { t0 += __shfl_xor_sync(0xffffffffu, t0, off); t1 += __shfl_xor_sync(0xffffffffu, t1, off); }
    if (tid < 2) base[tid] = 0;
    __syncthreads();
    if (lane == 0) { atomicAdd(&base[0], t0); atomicAdd(&base[1], t1); }       // integer sums: order-independent
    const unsigned br = __ballot_sync(0xffffffffu, fr), bi = __ballot_sync(0xffffffffu, fi);
    if (lane == 0) { warp_tot[0][warp] = __popc(br); warp_tot[1][warp] = __popc(bi); }
    __syncthreads();
    int pr = base[0], pi = base[1];
    for (int w = 0; w < warp; ++w) { pr += warp_tot[0][w]; pi += warp_tot[1][w]; }
    rr = pr + __popc(br & ((1u << lane) - 1u));
    ri = pi + __popc(bi & ((1u << lane) - 1u));
  }
  row_r[tid] = fr ? rr : -1;
  row_i[tid] = fi ? ri : -1;
  __syncthreads();
  // ---- the tile's [TILE, dim] span, 128-bit accesses
  const int d4 = s.dim >> 2;
  const int e0 = blockIdx.x * TILE;
  const int rows = min(TILE, s.num_envs - e0);
  float4* buf4 = reinterpret_cast<float4*>(a.buffer) + (size_t)e0 * d4;
  float4* cor4 = reinterpret_cast<float4*>(s.correlated) + (size_t)e0 * d4;
  for (int i = tid; i < rows * d4; i += TILE) {
    const int el = i / d4, q = i - el * d4;
    float4 x = buf4[i];
    if (s.on_reset.enabled) {
      float4 cn;
      const int row = row_r[el];
      if (row >= 0) {
        const float4 d = a.draws_reset ? __ldg(reinterpret_cast<const float4*>(a.draws_reset) + (size_t)row * d4 + q)
                                       : draw4(a, s.on_reset.distribution, (unsigned long long)(e0 + el) * d4 + q, 0u);
        cn = shape4(s.on_reset, d);
        cor4[i] = cn;
      } else {
        cn = cor4[i];
      }
      x = op4(s.on_reset.operation, x, cn);
    }
    const int rowi = row_i[el];
    if (rowi >= 0) {
      const float4 d = a.draws_interval ? __ldg(reinterpret_cast<const float4*>(a.draws_interval) + (size_t)rowi * d4 + q)
                                        : draw4(a, s.on_interval.distribution, (unsigned long long)(e0 + el) * d4 + q, 1u);
      x = op4(s.on_interval.operation, x, shape4(s.on_interval, d));
    }
    buf4[i] = x;
  }
}

}  // namespace noise
}  // namespace lm

using namespace lm;
using namespace lm::noise;

struct LmNoise {
  int device;
  int* counts;
  int counts_tiles;
};

extern "C" {

LmNoise* lm_noise_create(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { lm_set_error("lm_noise_create: %s", cudaGetErrorString(e)); return nullptr; }
  if (prop.major != 10) { lm_set_error("lm_noise_create: library is sm_100a only (device sm_%d%d)", prop.major, prop.minor); return nullptr; }
  cudaSetDevice(device);
  LmNoise* h = new LmNoise();
  h->device = device; h->counts = nullptr; h->counts_tiles = 0;
  return h;
}

void lm_noise_destroy(LmNoise* h) {
  if (!h) return;
  cudaFree(h->counts);
  delete h;
}

int lm_noise_apply(LmNoise* h, const LmNoiseState* state, float* buffer, const int64_t* reset_buf, const float* draws_reset,
                   const float* draws_interval, int32_t compact_draws, uint64_t seed, uint64_t step, void* stream) {
  if (!h || !state || !buffer || !reset_buf) return lm_fail(LM_ERR_BAD_ARG, "lm_noise_apply: NULL argument");
  const LmNoiseState& s = *state;
  if (!s.counter || s.num_envs <= 0 || s.dim <= 0 || (s.dim & 3)) return lm_fail(LM_ERR_BAD_ARG, "lm_noise_apply: counter NULL or dim not a multiple of 4");
  if (s.on_reset.enabled && !s.correlated) return lm_fail(LM_ERR_BAD_ARG, "lm_noise_apply: on_reset noise needs the correlated buffer");
  if ((((uintptr_t)buffer) | ((uintptr_t)s.correlated) | ((uintptr_t)draws_reset) | ((uintptr_t)draws_interval)) & 15)
    return lm_fail(LM_ERR_ALIGNMENT, "lm_noise_apply: buffer / correlated / draws must be 16-byte aligned");
  for (const LmNoiseSpec* sp : {&s.on_reset, &s.on_interval}) {
    if (!sp->enabled) continue;
    if (sp->operation != LM_NOISE_ADDITIVE && sp->operation != LM_NOISE_SCALING) return lm_fail(LM_ERR_UNSUPPORTED, "lm_noise_apply: unknown operation");
    if (sp->distribution < LM_DIST_GAUSSIAN || sp->distribution > LM_DIST_LOGUNIFORM) return lm_fail(LM_ERR_UNSUPPORTED, "lm_noise_apply: unknown distribution");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = (s.num_envs + TILE - 1) / TILE;
  Args a;
  a.s = s; a.buffer = buffer; a.reset_buf = reinterpret_cast<const long long*>(reset_buf);
  a.draws_reset = draws_reset; a.draws_interval = draws_interval;
  a.compact = (compact_draws && (draws_reset || draws_interval)) ? 1 : 0;
  a.seed = seed; a.step = step; a.counts = nullptr;
  if (a.compact) {
    if (h->counts_tiles < tiles) {
      cudaFree(h->counts);
      if (cudaMalloc(&h->counts, (size_t)tiles * 2 * sizeof(int)) != cudaSuccess) { h->counts = nullptr; h->counts_tiles = 0; return lm_fail((int)cudaErrorMemoryAllocation, "lm_noise_apply: cudaMalloc failed"); }
      h->counts_tiles = tiles;
    }
    a.counts = h->counts;
    k_noise_count<<<tiles, TILE, 0, st>>>(a);
  }
  k_noise_apply<<<tiles, TILE, 0, st>>>(a);
  return lm_check_launch("lm_noise_apply");
}

}  // extern "C"
