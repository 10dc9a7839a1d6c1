// GNN limb-graph policy forward on 5th-gen tensor cores (tcgen05 + TMEM): bf16 operands, fp32 accumulation.
// Same network and packing order as lm_gnn.cu (scripts/graph_model_orebot_ov.py:11-241); logits within 1e-2 of fp32.
//
// Mapping.  A tile is 128 envs = the 128 rows (TMEM lanes) of every MMA; the 13 node feature vectors of the tile live
// in shared memory as 13 bf16 A operands h[n] = [128 envs x 32] (UMMA canonical K-major).  The edge MLP of edge (i <- j)
//     m_ij = ELU(W2 ELU(W1a h_i + W1b h_j + b1) + b2)                      (:53-59; i = target)
// is two accumulating MMAs into one 32-column TMEM slot (A = h[i] against W1a, then A = h[j] against W1b: the
// concatenation [h_i ; h_j] of :55 is never materialised and no row is gathered), an epilogue (+b1, ELU, bf16) that
// writes z_ij as the next A operand, and one MMA against W2 into the same slot.  The aggregation h_i' = max_j m_ij
// (:46) is taken on the pre-activations, ELU being monotone: 13 instead of 24 second ELUs per layer.
//
// The limb graph is a star of four 3-joint chains (0 - k - k+4 - k+8, k = 1..4; :142-159).  A "leg" is the 6 directed
// edges of one chain.  After stage 1 of a leg has completed, nodes k, k+4, k+8 are not read again in this layer, so
// their h' is written IN PLACE; node 0's running maximum stays in registers until the fourth leg.  Per layer the 16
// epilogue warps (4 TMEM lane groups x 4 column quarters: a thread owns 8 of the 32 hidden channels of one env) run
//     E1(leg1) E1(leg2) E2(leg1) E2(leg2) E1(leg3) E1(leg4) E2(leg3) E2(leg4)
// over two sets of 6 TMEM slots / 6 z operands, and the control thread issues the MMAs of a stage as soon as its inputs
// exist, so every wait on the tensor pipe has another leg's epilogue in front of it.
// The input layers (:97-126; K = 16 and K = 4) and the two heads (:215-241) are fp32 CUDA-core work on the same thread
// mapping.  Activations never touch HBM: 256 B read and 52 B written per env.
#include <cuda_bf16.h>

#include "lm_common.cuh"
#include "lm_gnn.h"
#include "lm_host.h"
#include "lm_tc.cuh"

namespace lm {
namespace gnn_tc {
using namespace tc;

constexpr int TM = 128;
constexpr int E_WARPS = 16;
constexpr int E_THREADS = E_WARPS * 32;
constexpr int THREADS = E_THREADS + 32;
constexpr int H = 32;
constexpr int NODES = 13;
constexpr int LEG = 6;                               // directed edges per leg
constexpr int OPB = operand_bytes(TM, H);            // one [128 x 32] bf16 A operand
constexpr int WB = operand_bytes(H, H);              // one [32 x 32] bf16 B operand

// parameter blob (global, produced by lm_gnn_tc_pack; mirrored 1:1 at the start of shared memory)
constexpr int WL_OFF = 0;                            // layer l: W1a, W1b, W2 at (3 l + m) * WB
constexpr int F_OFF = 9 * WB;                        // fp32 parameters, offsets in floats:
constexpr int F_WIN1 = 0;                            //   input_layer1 W^T [16][32]
constexpr int F_WIN2 = F_WIN1 + 16 * H;              //   input_layer2 W^T [4][32]
constexpr int F_BIN1 = F_WIN2 + 4 * H;
constexpr int F_BIN2 = F_BIN1 + H;
constexpr int F_BL = F_BIN2 + H;                     //   per layer b1[32], b2[32]
constexpr int F_WACT = F_BL + 3 * 2 * H;
constexpr int F_WVAL = F_WACT + H;
constexpr int F_BHEAD = F_WVAL + H;                  //   b_action, b_value, 2 pad
constexpr int F_COUNT = F_BHEAD + 4;
constexpr int BLOB_BYTES = F_OFF + F_COUNT * 4;
static_assert(BLOB_BYTES % 16 == 0 && F_OFF % 16 == 0, "align");

constexpr int H_OFF = BLOB_BYTES;                    // 13 node operands
constexpr int Z_OFF = H_OFF + NODES * OPB;           // 2 sets x 6 z operands
constexpr int SMEM_BYTES = Z_OFF + 2 * LEG * OPB;
constexpr int OBS_STRIDE = 68;                       // floats per staged observation row (+4: conflict-free row reads)
constexpr int STAGE_OFF = Z_OFF + LEG * OPB;         // observation staging lives in z set 1 (idle between tiles)
static_assert(TM * OBS_STRIDE * 4 <= LEG * OPB, "obs staging fits one z set");
static_assert(SMEM_BYTES + 256 <= 232448, "exceeds 227 KB");
constexpr uint32_t TMEM_COLS = 512;                  // 12 slots x 32 columns used

struct Args {
  const uint8_t* blob;
  const float* obs;
  float* mu;
  float* value;
  int n;
  int num_tiles;
  long long* timeline;
};

__device__ __forceinline__ void stamp(const Args& a, int k) {
  if (a.timeline != nullptr && blockIdx.x == 0) a.timeline[k] = clock64();
}
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(E_THREADS) : "memory"); }

// ---------------------------------------------------------------- control thread: MMA issue
// one K = 32 product (two K/16 steps) accumulated into `acc`
__device__ __forceinline__ void mma_k32(uint32_t acc, uint32_t a_saddr, uint32_t b_saddr, uint32_t accumulate) {
  constexpr uint32_t idesc = instr_desc(TM, H);
  umma_bf16(acc, smem_desc(a_saddr, lbo_bytes(TM)), smem_desc(b_saddr, lbo_bytes(H)), idesc, accumulate);
  umma_bf16(acc, smem_desc(a_saddr + 2 * lbo_bytes(TM), lbo_bytes(TM)), smem_desc(b_saddr + 2 * lbo_bytes(H), lbo_bytes(H)), idesc, 1u);
}
// edges of leg k, as (target, source): e0 (k<-0) e1 (k<-k+4) e2 (k+4<-k) e3 (k+4<-k+8) e4 (k+8<-k+4) e5 (0<-k)
__device__ __forceinline__ void issue_stage1(uint32_t tmem_set, uint32_t sH, uint32_t sW1a, uint32_t sW1b, int k) {
  const int tgt[LEG] = {k, k, k + 4, k + 4, k + 8, 0};
  const int src[LEG] = {0, k + 4, k, k + 8, k + 4, k};
#pragma unroll
  for (int e = 0; e < LEG; ++e) {
    mma_k32(tmem_set + e * H, sH + tgt[e] * OPB, sW1a, 0u);
    mma_k32(tmem_set + e * H, sH + src[e] * OPB, sW1b, 1u);
  }
}
// leg 1 of the NEXT layer, issued while the current layer is still in E2(leg 4): everything that does not read h[0]
// (h'[1], h'[5], h'[9] were written by E2(leg 1)); e0 and e5 get their node-0 halves from issue_stage1_leg1_late.
__device__ __forceinline__ void issue_stage1_leg1_early(uint32_t tmem_set, uint32_t sH, uint32_t sW1a, uint32_t sW1b) {
  mma_k32(tmem_set + 0 * H, sH + 1 * OPB, sW1a, 0u);                                                  // e0 (1 <- 0): target half
  mma_k32(tmem_set + 1 * H, sH + 1 * OPB, sW1a, 0u); mma_k32(tmem_set + 1 * H, sH + 5 * OPB, sW1b, 1u);   // e1 (1 <- 5)
  mma_k32(tmem_set + 2 * H, sH + 5 * OPB, sW1a, 0u); mma_k32(tmem_set + 2 * H, sH + 1 * OPB, sW1b, 1u);   // e2 (5 <- 1)
  mma_k32(tmem_set + 3 * H, sH + 5 * OPB, sW1a, 0u); mma_k32(tmem_set + 3 * H, sH + 9 * OPB, sW1b, 1u);   // e3 (5 <- 9)
  mma_k32(tmem_set + 4 * H, sH + 9 * OPB, sW1a, 0u); mma_k32(tmem_set + 4 * H, sH + 5 * OPB, sW1b, 1u);   // e4 (9 <- 5)
  mma_k32(tmem_set + 5 * H, sH + 1 * OPB, sW1b, 0u);                                                  // e5 (0 <- 1): source half
}
__device__ __forceinline__ void issue_stage1_leg1_late(uint32_t tmem_set, uint32_t sH, uint32_t sW1a, uint32_t sW1b) {
  mma_k32(tmem_set + 0 * H, sH, sW1b, 1u);
  mma_k32(tmem_set + 5 * H, sH, sW1a, 1u);
}
__device__ __forceinline__ void issue_stage2(uint32_t tmem_set, uint32_t sZset, uint32_t sW2) {
#pragma unroll
  for (int e = 0; e < LEG; ++e) mma_k32(tmem_set + e * H, sZset + e * OPB, sW2, 0u);
}

// ---------------------------------------------------------------- epilogue helpers
__device__ __forceinline__ void load_bias8(const float* p, float b[8]) {
  const float4 lo = *reinterpret_cast<const float4*>(p), hi = *reinterpret_cast<const float4*>(p + 4);
  b[0] = lo.x; b[1] = lo.y; b[2] = lo.z; b[3] = lo.w; b[4] = hi.x; b[5] = hi.y; b[6] = hi.z; b[7] = hi.w;
}
// 8 accumulator columns + bias -> ELU -> 8 bf16 = one 16-byte K chunk
#ifndef LM_GNN_FMA_PAIRS   // of the 4 ELU pairs of a chunk, how many take the FMA-pipe exp (tuned on B200)
#define LM_GNN_FMA_PAIRS 1
#endif
// Node features and edge activations are stored SCALED by log2(e) (LM_GNN_SCALED, default; lm_tc.cuh elu_pair_scaled): the tensor-core
// blob carries log2(e) x the input-layer weights / biases and log2(e) x the graph-layer biases, the graph-layer weights are plain, so
// every accumulator + bias is the ex2 argument itself (7 instead of 8 instructions per ELU pair); the fp32 last layer un-scales.
#ifndef LM_GNN_SCALED
#define LM_GNN_SCALED 1
#endif
constexpr float kActScale = LM_GNN_SCALED ? 1.4426950408889634f : 1.0f;
__device__ __forceinline__ uint4 elu_chunk(const uint32_t r[8], const float b[8]) {
  constexpr int F = LM_GNN_FMA_PAIRS;
#if LM_GNN_SCALED
  return make_uint4(elu_pair_mix_scaled<F>(0, r[0], r[1], b[0], b[1]), elu_pair_mix_scaled<F>(1, r[2], r[3], b[2], b[3]),
                    elu_pair_mix_scaled<F>(2, r[4], r[5], b[4], b[5]), elu_pair_mix_scaled<F>(3, r[6], r[7], b[6], b[7]));
#else
  return make_uint4(elu_pair_mix<F>(0, r[0], r[1], b[0], b[1]), elu_pair_mix<F>(1, r[2], r[3], b[2], b[3]),
                    elu_pair_mix<F>(2, r[4], r[5], b[4], b[5]), elu_pair_mix<F>(3, r[6], r[7], b[6], b[7]));
#endif
}
__device__ __forceinline__ uint4 pack_chunk(const float v[8]) {
  return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ void max8(uint32_t a[8], const uint32_t b[8]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c] = __float_as_uint(fmaxf(__uint_as_float(a[c]), __uint_as_float(b[c])));
}
// fp32 ELU of 8 channels (last layer: the heads read fp32, not the bf16 operand)
__device__ __forceinline__ void elu8(const uint32_t r[8], const float b[8], float v[8]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = elu((__uint_as_float(r[c]) + b[c]) * (1.0f / kActScale));   // accumulator and bias carry the activation scale
}

// E1 of one leg: stage-1 accumulators (6 slots) -> z operands.  taddr = this thread's lane group / column quarter in slot 0 of the set.
__device__ __forceinline__ void epilogue_stage1(uint32_t taddr, const float b1[8], uint8_t* zdst /* + q*lbo + row*16 applied */,
                                                long long* dg = nullptr) {
  uint32_t r[LEG][8];
#pragma unroll
  for (int e = 0; e < 3; ++e) tmem_ld8(taddr + e * H, r[e]);
  tmem_wait();
  if (dg) dg[0] = clock64();
#pragma unroll
  for (int e = 3; e < LEG; ++e) tmem_ld8(taddr + e * H, r[e]);
#pragma unroll
  for (int e = 0; e < 3; ++e) *reinterpret_cast<uint4*>(zdst + e * OPB) = elu_chunk(r[e], b1);
  if (dg) dg[1] = clock64();
  tmem_wait();
  if (dg) dg[2] = clock64();
#pragma unroll
  for (int e = 3; e < LEG; ++e) *reinterpret_cast<uint4*>(zdst + e * OPB) = elu_chunk(r[e], b1);
  if (dg) dg[3] = clock64();
}

// E2 of leg `leg`: stage-2 accumulators -> max aggregation (pre-activation: ELU is monotone) -> nodes leg, leg+4, leg+8; node 0
// keeps a running max.  hdst = this thread's 16-byte chunk in h[leg].  Graph layers 1-2 write h' (bf16 operand) in place.  The last
// layer keeps fp32: the thread's 8-channel share of action_layer . h[n] goes into the same (now dead) 16-byte slot, where
// the head reduction collects the four column quarters, and vmax follows the per-channel maximum over the nodes.
__device__ __forceinline__ void epilogue_stage2(uint32_t taddr, const float b2[8], uint8_t* hdst, uint32_t m0[8], bool last,
                                                const float* __restrict__ wact_sm, float vmax[8]) {
  uint32_t r[LEG][8];
#pragma unroll
  for (int e = 0; e < LEG; ++e) tmem_ld8(taddr + e * H, r[e]);
  tmem_wait();
  max8(r[0], r[1]);          // node k    <- {0, k+4}
  max8(r[2], r[3]);          // node k+4  <- {k, k+8}
  max8(m0, r[5]);            // node 0    <- k
  if (!last) {
#pragma unroll
    for (int t = 0; t < 3; ++t) *reinterpret_cast<uint4*>(hdst + 4 * t * OPB) = elu_chunk(r[2 * t], b2);
  } else {
    float wact[8];
    load_bias8(wact_sm, wact);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      float v[8], s = 0.0f;
      elu8(r[2 * t], b2, v);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        s = fmaf(wact[c], v[c], s);
        vmax[c] = fmaxf(vmax[c], v[c]);
      }
      *reinterpret_cast<float*>(hdst + 4 * t * OPB) = s;
    }
  }
}

// coalesced 128-bit loads: thread owns float4 j*512 + tid of the [128 x 64] fp32 tile (row j*32 + tid/16, columns 4 (tid%16) ..)
__device__ __forceinline__ void obs_fetch(const Args& a, int row0, int rows, int tid, float4 v[4]) {
  const float* base = a.obs + (size_t)row0 * 64;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    v[j] = (j * 32 + (tid >> 4)) < rows ? ld_nc_f4(base + (size_t)(j * E_THREADS + tid) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void obs_stage(float* stage, int tid, const float4 v[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(stage + (j * 32 + (tid >> 4)) * OBS_STRIDE + 4 * (tid & 15)) = v[j];
}

// per-thread view of the epilogue side of the pipeline
struct EpiCtx {
  uint64_t *bar_h, *bar_s1, *bar_s2, *bar_z, *bar_e2;
  uint32_t t_set0, t_set1;      // TMEM address of this thread's lane group / column quarter in slot 0 of each set
  uint8_t *hmine, *z0, *z1;     // this thread's 16-byte K chunk in h[0] / z[set][0]
  const float* F;
  int q, lane;
  long long* tl;                // diagnostic timeline of this tile (thread 0 of CTA 0 only, else NULL)
};
__device__ __forceinline__ void tstamp(const EpiCtx& c, int k) {
  if (c.tl != nullptr) c.tl[k] = clock64();
}

// One graph layer on the epilogue side: E1(leg1) E1(leg2) E2(leg1) E2(leg2) E1(leg3) E1(leg4) E2(leg3) E2(leg4), then node 0.
// Rolled loops on purpose: one copy of each stage body keeps the kernel inside the instruction cache (the fully unrolled
// version lost 14 % of its issue slots to instruction fetch, ncu stall_no_inst).
__device__ __forceinline__ void run_layer(const EpiCtx& c, int layer, float vmax[8]) {
  const bool last = layer == 2;
  float b1[8], b2[8];
  load_bias8(c.F + F_BL + layer * 2 * H + 8 * c.q, b1);
  load_bias8(c.F + F_BL + layer * 2 * H + H + 8 * c.q, b2);
  const float* wact_sm = c.F + F_WACT + 8 * c.q;
  uint32_t m0[8];
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) m0[ch] = 0xff800000u;      // -inf
  const int tb = 8 + 16 * layer;
#pragma unroll 1
  for (int par = 0; par < 2; ++par) {
#pragma unroll
    for (int set = 0; set < 2; ++set) {
      mbar_wait(&c.bar_s1[set], par); tc_fence_after(); tstamp(c, tb + 8 * par + 2 * set);
      epilogue_stage1(c.t_set0 + set * LEG * H, b1, c.z0 + set * LEG * OPB, (c.tl && layer == 1 && par + set == 0) ? c.tl + 56 : nullptr);
      fence_proxy_async(); tc_fence_before(); __syncwarp();
      if (c.lane == 0) mbar_arrive(&c.bar_z[set]);
      tstamp(c, tb + 8 * par + 2 * set + 1);
    }
#pragma unroll
    for (int set = 0; set < 2; ++set) {
      const int leg = 1 + 2 * par + set;
      mbar_wait(&c.bar_s2[set], par); tc_fence_after(); tstamp(c, tb + 8 * par + 4 + 2 * set);
      epilogue_stage2(c.t_set0 + set * LEG * H, b2, c.hmine + leg * OPB, m0, last, wact_sm, vmax);
      fence_proxy_async(); tc_fence_before(); __syncwarp();
      if (c.lane == 0) mbar_arrive(&c.bar_e2[set]);
      tstamp(c, tb + 8 * par + 4 + 2 * set + 1);
    }
  }
  // node 0 <- {1, 2, 3, 4}: every stage-1 MMA of this layer has completed, h[0] may be overwritten
  if (!last) {
    *reinterpret_cast<uint4*>(c.hmine) = elu_chunk(m0, b2);
    fence_proxy_async();
    __syncwarp();
    if (c.lane == 0) mbar_arrive(c.bar_h);
  } else {
    float v[8], wv[8], sv = 0.0f;
    elu8(m0, b2, v);
    load_bias8(c.F + F_WVAL + 8 * c.q, wv);
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) sv = fmaf(wv[ch], fmaxf(vmax[ch], v[ch]), sv);
    *reinterpret_cast<float*>(c.hmine) = sv;                // value-head share, in the dead h[0] slot
  }
}

__global__ void __launch_bounds__(THREADS, 1) k_gnn_forward_tc(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // mbarriers.  count 1 (tcgen05.commit): bar_w (weight TMA), bar_s1[s] / bar_s2[s] (stage 1 / 2 of the leg in set s complete).
  // count 16 (one arrive per epilogue warp): bar_h (all 13 node operands of the next layer written), bar_z[s] (z operands of
  // set s written = its stage-1 accumulators drained), bar_e2[s] (stage-2 accumulators of set s drained, h' written).
  // Per layer every [s] barrier completes exactly twice (legs s+1 and s+3), so wait parities are 0 then 1 in every layer;
  // bar_h completes three times per tile and carries a running phase.
  __shared__ uint64_t bar_w, bar_h, bar_s1[2], bar_s2[2], bar_z[2], bar_e2[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* F = reinterpret_cast<const float*>(smem + F_OFF);

  pdl_launch_dependents();
  if (tid == 0) {
    mbar_init(&bar_w, 1); mbar_init(&bar_h, E_WARPS);
    for (int s = 0; s < 2; ++s) { mbar_init(&bar_s1[s], 1); mbar_init(&bar_s2[s], 1); mbar_init(&bar_z[s], E_WARPS); mbar_init(&bar_e2[s], E_WARPS); }
    fence_mbar_init();
  }
  if (warp == E_WARPS) tmem_alloc(&tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // (Dealing rows in 32-row groups so that the last round is a half tile was measured and dropped: a tile whose upper lane
  // groups are idle takes as long as a full one - the epilogue is bound by each warp's own dependent chain, not by throughput.)
  const int tile0 = blockIdx.x, stride = gridDim.x;
  const int n_loc = tile0 < a.num_tiles ? (a.num_tiles - tile0 + stride - 1) / stride : 0;
  if (tid == 0) stamp(a, 0);

  if (warp == E_WARPS) {
    // ================================================================== control warp
    if (elect_one()) {
      mbar_expect_tx(&bar_w, BLOB_BYTES);
      tma_load_1d(smem, a.blob, BLOB_BYTES, &bar_w);          // resident weights: one bulk copy per CTA
      mbar_wait(&bar_w, 0);
      const uint32_t sH = smem_u32(smem + H_OFF), sZ = smem_u32(smem + Z_OFF), sW = smem_u32(smem + WL_OFF);
      const uint32_t set0 = tmem, set1 = tmem + LEG * H;
      uint32_t ph_h = 0;
      for (int it = 0; it < n_loc; ++it) {
#pragma unroll 1
        for (int layer = 0; layer < 3; ++layer) {
          const uint32_t sW1a = sW + (3 * layer) * WB, sW1b = sW1a + WB, sW2 = sW1b + WB;
          mbar_wait(&bar_h, ph_h); ph_h ^= 1;
          tc_fence_after();
          long long* tl = (a.timeline != nullptr && blockIdx.x == 0 && it < 7 && layer == 0) ? a.timeline + 8 + 64 * it : nullptr;
          if (tl) tl[6] = clock64();
          if (layer == 0) issue_stage1(set0, sH, sW1a, sW1b, 1);
          else issue_stage1_leg1_late(set0, sH, sW1a, sW1b);
          if (tl) tl[7] = clock64();
          umma_commit(&bar_s1[0]);
          issue_stage1(set1, sH, sW1a, sW1b, 2); umma_commit(&bar_s1[1]);
          mbar_wait(&bar_z[0], 0); tc_fence_after(); issue_stage2(set0, sZ, sW2); umma_commit(&bar_s2[0]);
          mbar_wait(&bar_z[1], 0); tc_fence_after(); issue_stage2(set1, sZ + LEG * OPB, sW2); umma_commit(&bar_s2[1]);
          mbar_wait(&bar_e2[0], 0); tc_fence_after(); issue_stage1(set0, sH, sW1a, sW1b, 3); umma_commit(&bar_s1[0]);
          mbar_wait(&bar_e2[1], 0); tc_fence_after(); issue_stage1(set1, sH, sW1a, sW1b, 4); umma_commit(&bar_s1[1]);
          mbar_wait(&bar_z[0], 1); tc_fence_after(); issue_stage2(set0, sZ, sW2); umma_commit(&bar_s2[0]);
          mbar_wait(&bar_z[1], 1); tc_fence_after(); issue_stage2(set1, sZ + LEG * OPB, sW2); umma_commit(&bar_s2[1]);
          if (layer < 2) {                                   // E2(leg 3) done: set 0 is free, h'[1], h'[5], h'[9] exist
            mbar_wait(&bar_e2[0], 1); tc_fence_after();
            issue_stage1_leg1_early(set0, sH, sW1a + 3 * WB, sW1b + 3 * WB);
          }
        }
      }
    }
  } else {
    // ================================================================== epilogue warps
    const int g = warp & 3, q = warp >> 2;
    const int row = g * 32 + lane;
    const uint32_t t_set0 = tmem + ((uint32_t)(g * 32) << 16) + (uint32_t)(q * 8), t_set1 = t_set0 + LEG * H;
    const int my_off = q * lbo_bytes(TM) + row * 16;        // this thread's 16-byte K chunk inside any [128 x 32] operand
    uint8_t* hmine = smem + H_OFF + my_off;
    uint8_t* z0 = smem + Z_OFF + my_off;
    uint8_t* z1 = z0 + LEG * OPB;
    float* stage = reinterpret_cast<float*>(smem + STAGE_OFF);
    const bool st = tid == 0;
    EpiCtx c;
    c.bar_h = &bar_h; c.bar_s1 = bar_s1; c.bar_s2 = bar_s2; c.bar_z = bar_z; c.bar_e2 = bar_e2;
    c.t_set0 = t_set0; c.t_set1 = t_set1; c.hmine = hmine; c.z0 = z0; c.z1 = z1; c.F = F; c.q = q; c.lane = lane;
    pdl_wait();                                              // the observations are the predecessor kernel's output
    mbar_wait(&bar_w, 0);                                    // fp32 parameters live in the weight blob
    for (int it = 0; it < n_loc; ++it) {
      const int row0 = (tile0 + it * stride) * TM, rows = min(TM, a.n - row0);
      c.tl = (st && a.timeline != nullptr && blockIdx.x == 0 && it < 7) ? a.timeline + 8 + 64 * it : nullptr;
      tstamp(c, 0);
      // ---- observations -> padded fp32 staging (the first tile's here; later tiles' were fetched during the previous tile's heads)
      if (it == 0) {
        float4 v[4];
        obs_fetch(a, row0, rows, tid, v);
        obs_stage(stage, tid, v);
      }
      if (it + 1 < n_loc && tid < 256) {                     // the tile after this one -> L2
        const char* nxt = reinterpret_cast<const char*>(a.obs + (size_t)(row0 + stride * TM) * 64) + tid * 128;
        if (nxt < reinterpret_cast<const char*>(a.obs + (size_t)a.n * 64)) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
      }
      epi_sync();
      // ---- input layers (:97-126): node 0 <- Linear(16,32)(obs[0:16]); node 1+j <- Linear(4,32)(jp, jv, a_cur, a_last of joint pi(j))
      {
        const float* x = stage + row * OBS_STRIDE;
        float acc[8];
        load_bias8(F + F_BIN1 + 8 * q, acc);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const float4 xv = *reinterpret_cast<const float4*>(x + 4 * k4);
          const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            float w[8];
            load_bias8(F + F_WIN1 + (4 * k4 + kk) * H + 8 * q, w);
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = fmaf(w[c], xs[kk], acc[c]);
          }
        }
        *reinterpret_cast<uint4*>(hmine) = pack_chunk(acc);
        float w0[8], w1[8], w2[8], w3[8], bj[8];
        load_bias8(F + F_WIN2 + 0 * H + 8 * q, w0); load_bias8(F + F_WIN2 + 1 * H + 8 * q, w1);
        load_bias8(F + F_WIN2 + 2 * H + 8 * q, w2); load_bias8(F + F_WIN2 + 3 * H + 8 * q, w3);
        load_bias8(F + F_BIN2 + 8 * q, bj);
        // joint pi's four features sit at obs[16 + pi], [28 + pi], [40 + pi], [52 + pi]: four 128-bit loads serve 4 joints
#pragma unroll
        for (int p4 = 0; p4 < 3; ++p4) {
          const float4 fa = *reinterpret_cast<const float4*>(x + 16 + 4 * p4), fb = *reinterpret_cast<const float4*>(x + 28 + 4 * p4);
          const float4 fc = *reinterpret_cast<const float4*>(x + 40 + 4 * p4), fd = *reinterpret_cast<const float4*>(x + 52 + 4 * p4);
          const float f0[4] = {fa.x, fa.y, fa.z, fa.w}, f1[4] = {fb.x, fb.y, fb.z, fb.w};
          const float f2[4] = {fc.x, fc.y, fc.z, fc.w}, f3[4] = {fd.x, fd.y, fd.z, fd.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int pi = 4 * p4 + i;
            const int j = pi < 4 ? pi : ((pi & 1) ? 8 + (pi - 5) / 2 : 4 + (pi - 4) / 2);   // inverse of pi(j) = (0,1,2,3, 4,6,8,10, 5,7,9,11)
            float y[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) y[c] = fmaf(w3[c], f3[i], fmaf(w2[c], f2[i], fmaf(w1[c], f1[i], fmaf(w0[c], f0[i], bj[c]))));
            *reinterpret_cast<uint4*>(hmine + (1 + j) * OPB) = pack_chunk(y);
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_h);
      tstamp(c, 1);

      float vmax[8];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) vmax[ch] = -INFINITY;
#pragma unroll 1
      for (int layer = 0; layer < 3; ++layer) {
        run_layer(c, layer, vmax);
        tstamp(c, 2 + layer);
      }
      // ---- heads (:215-241): action k = Wa . h[1+k] + ba; value = Wv . max_n h[n] + bv.  The four column quarters left
      // their fp32 shares in their own 16-byte slots of the (dead) node operands; 512 threads sum them, coalesced stores.
      float4 nobs[4];
      const int nrow0 = row0 + stride * TM;
      if (it + 1 < n_loc) obs_fetch(a, nrow0, min(TM, a.n - nrow0), tid, nobs);   // lands during the head reduction
      epi_sync();
      {
        const float ba = F[F_BHEAD], bv = F[F_BHEAD + 1];
        const uint8_t* hb = smem + H_OFF;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int idx = j * E_THREADS + tid, r = idx / 12, k = idx - 12 * r;
          const uint8_t* p0 = hb + (1 + k) * OPB + r * 16;
          const float s = ((*reinterpret_cast<const float*>(p0) + *reinterpret_cast<const float*>(p0 + lbo_bytes(TM))) +
                           *reinterpret_cast<const float*>(p0 + 2 * lbo_bytes(TM))) + *reinterpret_cast<const float*>(p0 + 3 * lbo_bytes(TM));
          if (r < rows) a.mu[(size_t)row0 * 12 + idx] = s + ba;
        }
        if (tid < TM) {
          const uint8_t* p0 = hb + tid * 16;
          const float s = ((*reinterpret_cast<const float*>(p0) + *reinterpret_cast<const float*>(p0 + lbo_bytes(TM))) +
                           *reinterpret_cast<const float*>(p0 + 2 * lbo_bytes(TM))) + *reinterpret_cast<const float*>(p0 + 3 * lbo_bytes(TM));
          if (tid < rows) a.value[row0 + tid] = s + bv;
        }
      }
      if (it + 1 < n_loc) obs_stage(stage, tid, nobs);        // z set 1 is idle: stage 2 of leg 4 completed before E2(leg 4)
      tstamp(c, 5);
      // the next tile's staging barrier (epi_sync) orders these reads before the input layer rewrites the node operands
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) stamp(a, 1);
  if (warp == E_WARPS) tmem_dealloc(tmem, TMEM_COLS);
}

// nn.Linear W [32, ld], columns [col0, col0 + 32) -> bf16 canonical B operand [32 x 32]
__global__ void k_pack_w(const float* __restrict__ w, uint8_t* __restrict__ dst, int ld, int col0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * H) return;
  const int r = i / H, k = i - r * H;
  *reinterpret_cast<__nv_bfloat16*>(dst + elem_offset(H, r, k)) = __float2bfloat16_rn(w[(size_t)r * ld + col0 + k]);
}
// nn.Linear W [32, K] -> W^T [K][32] fp32
__global__ void k_pack_wt(const float* __restrict__ w, float* __restrict__ dst, int K, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * K) return;
  const int c = i / K, k = i - c * K;
  dst[k * H + c] = w[i] * scale;
}
__global__ void k_copy(const float* __restrict__ src, float* __restrict__ dst, int n, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * scale;
}

}  // namespace gnn_tc
}  // namespace lm

using namespace lm;
using namespace lm::gnn_tc;

int lm_gnn_tc_init(LmGnn* g) {
  if (cudaMalloc(&g->tc_blob, BLOB_BYTES) != cudaSuccess) return lm_fail((int)cudaErrorMemoryAllocation, "lm_gnn_create: cudaMalloc failed");
  cudaMemset(g->tc_blob, 0, BLOB_BYTES);
  cudaError_t e = cudaFuncSetAttribute(k_gnn_forward_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) { cudaFree(g->tc_blob); g->tc_blob = nullptr; return lm_fail((int)e, "lm_gnn_create: smem attribute: %s", cudaGetErrorString(e)); }
  g->tc_fresh_pack = true;
  return LM_OK;
}

void lm_gnn_tc_release(LmGnn* g) {
  if (g->tc_blob) cudaFree(g->tc_blob);
  g->tc_blob = nullptr;
}

int lm_gnn_tc_pack(LmGnn* g, const float* const* w, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* B = g->tc_blob;
  float* Fp = reinterpret_cast<float*>(B + F_OFF);
  const float sc = kActScale;      // activations (and therefore these fp32 parameters) carry the activation scale, see elu_chunk
  k_pack_wt<<<(H * 16 + 127) / 128, 128, 0, s>>>(w[0], Fp + F_WIN1, 16, sc);
  k_copy<<<1, 32, 0, s>>>(w[1], Fp + F_BIN1, H, sc);
  k_pack_wt<<<1, 128, 0, s>>>(w[2], Fp + F_WIN2, 4, sc);
  k_copy<<<1, 32, 0, s>>>(w[3], Fp + F_BIN2, H, sc);
  for (int l = 0; l < 3; ++l) {
    k_pack_w<<<(H * H + 127) / 128, 128, 0, s>>>(w[4 + 4 * l], B + WL_OFF + (3 * l + 0) * WB, 2 * H, 0);   // W1a: columns of h_i (target)
    k_pack_w<<<(H * H + 127) / 128, 128, 0, s>>>(w[4 + 4 * l], B + WL_OFF + (3 * l + 1) * WB, 2 * H, H);   // W1b: columns of h_j (source)
    k_copy<<<1, 32, 0, s>>>(w[5 + 4 * l], Fp + F_BL + l * 2 * H, H, sc);
    k_pack_w<<<(H * H + 127) / 128, 128, 0, s>>>(w[6 + 4 * l], B + WL_OFF + (3 * l + 2) * WB, H, 0);
    k_copy<<<1, 32, 0, s>>>(w[7 + 4 * l], Fp + F_BL + l * 2 * H + H, H, sc);
  }
  k_copy<<<1, 32, 0, s>>>(w[16], Fp + F_WACT, H, 1.0f);
  k_copy<<<1, 32, 0, s>>>(w[17], Fp + F_BHEAD, 1, 1.0f);
  k_copy<<<1, 32, 0, s>>>(w[18], Fp + F_WVAL, H, 1.0f);
  k_copy<<<1, 32, 0, s>>>(w[19], Fp + F_BHEAD + 1, 1, 1.0f);
  g->tc_fresh_pack = true;
  return lm_check_launch("lm_gnn_pack (tensor-core blob)");
}

extern "C" {

int lm_gnn_forward_bf16(LmGnn* g, const float* obs, float* mu, float* value, int32_t n, void* stream) {
  if (!g || !obs || !mu || !value || n <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_gnn_forward_bf16: bad argument");
  if (!g->ready) return lm_fail(LM_ERR_BAD_ARG, "lm_gnn_forward_bf16: weights not packed (call lm_gnn_pack)");
  if (((uintptr_t)obs) & 15) return lm_fail(LM_ERR_ALIGNMENT, "lm_gnn_forward_bf16: obs must be 16-byte aligned");
  Args a;
  a.blob = g->tc_blob; a.obs = obs; a.mu = mu; a.value = value; a.n = n; a.num_tiles = (n + TM - 1) / TM; a.timeline = g->timeline;
  const int grid = a.num_tiles < g->sm_count ? a.num_tiles : g->sm_count;
  const bool pdl = !g->tc_fresh_pack;    // the resident-weight TMA precedes pdl_wait(): only while the blob is not being rewritten upstream
  g->tc_fresh_pack = false;
  cudaError_t le = lm_launch(k_gnn_forward_tc, dim3(grid), dim3(THREADS), SMEM_BYTES, (cudaStream_t)stream, pdl, a);
  if (le != cudaSuccess) return lm_fail((int)le, "lm_gnn_forward_bf16: %s", cudaGetErrorString(le));
  return lm_check_launch("lm_gnn_forward_bf16");
}

int lm_gnn_set_timeline(LmGnn* g, long long* stamps_dev) {
  if (!g) return lm_fail(LM_ERR_BAD_ARG, "lm_gnn_set_timeline: NULL");
  g->timeline = stamps_dev;
  return LM_OK;
}

}  // extern "C"
