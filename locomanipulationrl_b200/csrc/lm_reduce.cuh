// Grid-wide reductions by the last-CTA pattern, shared by the caller-side kernels (lm_rollout.cu, lm_ppo.cu, lm_task_cc.cu):
// every CTA publishes its partial sums as doubles, takes a ticket, and the CTA that draws the last ticket sums the
// partials in a fixed order (deterministic run to run) and finishes the computation.
#pragma once
#include "lm_common.cuh"

namespace lm {

// Returns true in exactly one CTA, after every CTA's global writes made before the call are visible to it.  Resets the ticket
// for the next launch on the stream.  All threads of the CTA must call it.
__device__ __forceinline__ bool last_cta(unsigned int* ticket) {
  __shared__ unsigned int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
    if (s_last) { *ticket = 0u; __threadfence(); }
  }
  __syncthreads();
  return s_last != 0u;
}

// CTA-level sum of kN per-thread floats, as doubles: out[i] (i < kN) valid in threads tid < kN after the call.
// sm: [kThreads / 32][kN] doubles of shared memory.
template <int kN, int kThreads>
__device__ __forceinline__ double cta_sum(const float (&v)[kN], double (*sm)[kN]) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int i = 0; i < kN; ++i) {
    double d = (double)v[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
    if (lane == 0) sm[warp][i] = d;
  }
  __syncthreads();
  double d = 0.0;
  if (tid < kN) {
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) d += sm[w][tid];
  }
  return d;
}

// Sum over the grid's partials[cta][kN] by the last CTA: kSub subsets of CTAs x kN values, the loads of a thread independent of
// each other; result for value j in threads tid < kN.  part: [kSub][kN] doubles of shared memory; kSub * kN <= blockDim.x.
template <int kN, int kSub>
__device__ __forceinline__ double grid_sum(const double* partials, double (*part)[kN]) {
  const int tid = threadIdx.x, sub = tid / kN, j = tid - sub * kN;
  if (sub < kSub) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int c = sub;
    for (; c + 3 * kSub < (int)gridDim.x; c += 4 * kSub) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] += __ldcg(partials + (size_t)(c + kSub * u) * kN + j);
    }
    for (; c < (int)gridDim.x; c += kSub) acc[0] += __ldcg(partials + (size_t)c * kN + j);
    part[sub][j] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  }
  __syncthreads();
  double d = 0.0;
  if (tid < kN) {
#pragma unroll
    for (int u = 0; u < kSub; ++u) d += part[u][tid];
  }
  return d;
}

}  // namespace lm
