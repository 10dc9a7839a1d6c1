// One-shot all-reduce over NVLink peer memory (the PPO gradient all-reduce and the episode-stat reduction).
//
// The messages are tiny (58,637 + 12 + 8 floats = 229 KB for the MLP policy, a few dozen floats for the statistics), so the
// cost of a collective is pure latency: every rank PUSHES its contribution into a receive slot of every peer (plain stores
// through cudaIpc-mapped pointers travel over NVLink / NVSwitch), raises a per-source flag on every peer, waits for the P flags
// on its own side and sums the P slots locally in rank order.  One NVLink one-way latency instead of NCCL's launch + protocol
// cost, and — the point of the exercise — phase A (produce + push) is a device function, so a kernel that PRODUCES the
// values (k_wgrad_reduce: the sum of the per-CTA weight-gradient partials) pushes them as it computes them: no second launch,
// no pass over HBM between the compute step and the collective.  Every rank adds the slots in the same order, so the result is
// bit-identical on all ranks (replicated parameters stay replicated).
//
// Receive memory per rank (one cudaMalloc region, one IPC handle):
//   flags  [LM_COMM_MAX_WORLD]             monotonically increasing epoch written by source rank r
//   data   [2 parities][world][slot_floats]
// Epoch e uses parity e & 1.  A rank cannot run two epochs ahead of a peer (it needs the peer's flag of epoch e+1, which the
// peer raises only after it has finished reading epoch e), so two parities suffice.  The epoch counter lives in device memory
// and is advanced by the kernel itself: the launch is CUDA-graph capturable.
#pragma once
#include <stdint.h>

#include "lm_common.cuh"

namespace lm {
namespace comm {

constexpr int MAX_WORLD = 8;
constexpr unsigned long long SPIN_TIMEOUT_NS = 4000000000ull;      // a missing peer must not hang the GPU: 4 s, then an error word

struct Ctx {
  int rank, world;
  long long slot_floats;
  float* peer_data[MAX_WORLD];          // peer p's data region, mapped into this process (p == rank: the local one)
  unsigned int* peer_flags[MAX_WORLD];
  float* local_data;
  unsigned int* local_flags;
  unsigned int* ctl;                    // local: [0] arrive ticket, [1] exit ticket, [2] epoch, [3] error word
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// epoch of this launch (all CTAs read the same value: it is advanced by the LAST CTA to leave, see leave())
__device__ __forceinline__ unsigned int epoch_of(const Ctx& c) { return __ldcg(c.ctl + 2) + 1u; }

// phase A: element f of this rank's contribution -> slot `rank` of every peer
__device__ __forceinline__ void push(const Ctx& c, unsigned int epoch, long long f, float v) {
  const long long off = ((long long)(epoch & 1u) * c.world + c.rank) * c.slot_floats + f;
#pragma unroll
  for (int p = 0; p < MAX_WORLD; ++p)
    if (p < c.world) c.peer_data[p][off] = v;
}

// All threads of every CTA call this after their pushes.  The last CTA to arrive raises this rank's flag on every peer; then
// every CTA waits until all sources have raised theirs on this rank.  Returns false after a time-out (error word set).
__device__ __forceinline__ bool signal_and_wait(const Ctx& c, unsigned int epoch) {
  __shared__ int s_ok;
  __threadfence_system();                       // this thread's pushes are ordered before the flag store that follows the ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(c.ctl + 0, 1u);
    if (t == gridDim.x - 1) {
      __threadfence_system();
      c.ctl[0] = 0u;
      for (int p = 0; p < c.world; ++p) st_release_sys(c.peer_flags[p] + c.rank, epoch);
    }
  }
  if (threadIdx.x < 32) {
    bool ok = true;
    if ((int)threadIdx.x < c.world) {
      const unsigned long long t0 = gtimer();
      // epochs only grow; `>=` in wrap-around arithmetic
      while ((int)(ld_acquire_sys(c.local_flags + threadIdx.x) - epoch) < 0) {
        if (gtimer() - t0 > SPIN_TIMEOUT_NS) { ok = false; atomicExch(c.ctl + 3, 1u + threadIdx.x); break; }
        __nanosleep(64);
      }
    }
    ok = __all_sync(0xffffffffu, ok);
    if (threadIdx.x == 0) s_ok = ok ? 1 : 0;
  }
  __syncthreads();
  return s_ok != 0;
}

// phase B: sum over the sources in rank order (the same order on every rank)
__device__ __forceinline__ float sum(const Ctx& c, unsigned int epoch, long long f) {
  const float* base = c.local_data + (long long)(epoch & 1u) * c.world * c.slot_floats + f;
  float v[MAX_WORLD];
#pragma unroll
  for (int p = 0; p < MAX_WORLD; ++p) v[p] = p < c.world ? __ldcg(base + (long long)p * c.slot_floats) : 0.0f;
  float s = v[0];
#pragma unroll
  for (int p = 1; p < MAX_WORLD; ++p)
    if (p < c.world) s += v[p];
  return s;
}

// Last thing every CTA does: the last CTA to leave advances the epoch for the next launch.
__device__ __forceinline__ void leave(const Ctx& c) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(c.ctl + 1, 1u);
    if (t == gridDim.x - 1) { c.ctl[1] = 0u; atomicAdd(c.ctl + 2, 1u); }
  }
}

}  // namespace comm
}  // namespace lm

// library-internal: the device-side context of a handle (lm_comm.cu)
struct LmComm;
const lm::comm::Ctx* lm_comm_ctx(const LmComm* c);
long long lm_comm_slot_floats(const LmComm* c);
int lm_comm_max_ctas(const LmComm* c);
