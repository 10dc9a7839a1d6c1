// MLP actor-critic forward on 5th-gen tensor cores (tcgen05 + TMEM), persistent, weights resident.
//
// Network (scripts/skrl_ppo_locomanipulation.py:33-42,57-61; rl_games actor_critic from
// cfg/train/JointLocomanipulationPPO.yaml:10-32):
//   x[64] -> Linear(64,256) ELU -> Linear(256,128) ELU -> Linear(128,64) ELU -> { mean Linear(64,12), value Linear(64,1) }
//
// One CTA per SM: 16 epilogue warps (4 TMEM lane groups x 4 column quarters) + 1 control warp whose lane 0 issues
// every tcgen05.mma.  All four weight matrices (bf16, 115 KB) are pulled into shared memory once per CTA by a TMA
// bulk copy and stay there; the CTA then walks 128-env tiles as a chunk-level pipeline:
//   * the epilogue of layer L (tcgen05.ld -> +bias, ELU, bf16 -> next A operand in shared memory) proceeds in 64-column
//     sub-steps; each finished sub-step is one K chunk of layer L+1, whose MMAs are issued at once, so the tensor
//     pipe runs underneath the epilogue instead of after it;
//   * every layer has its own TMEM accumulator columns (256 + 128 + 64 + 16 of the 512), so the next tile's layer-1
//     MMA is issued while the current tile is still in layers 2-4;
//   * the next tile's observations are fetched into registers one tile ahead (coalesced 128-bit loads), normalised/clamped
//     and written as the bf16 A0 operand inside the layer-1 epilogue's sub-steps (one float4 per sub-step).
// Activations never touch HBM; per env the kernel reads 256 B and writes 52 B.
#include <cuda_bf16.h>
#include <string.h>

#include "lm_common.cuh"
#include "lm_host.h"
#include "lm_mlp.h"
#include "lm_sink.cuh"
#include "lm_tc.cuh"

namespace lm {
namespace mlp {
using namespace tc;

#ifndef LM_MLP_FMA_PAIRS
#define LM_MLP_FMA_PAIRS 0
#endif
#ifndef LM_MLP_PIECES
#define LM_MLP_PIECES 2
#endif
constexpr int kPc = LM_MLP_PIECES;        // 16-column pieces per warp between two proxy fences (layers 1 and 2)
constexpr int TM = 128;                 // rows (envs) per tile == UMMA M
constexpr int E_WARPS = 16;               // epilogue warps: 4 TMEM lane groups x 4 column quarters
constexpr int E_THREADS = E_WARPS * 32;
constexpr int THREADS = E_THREADS + 32;   // + the control warp (MMA issue, weight TMA)
constexpr int D0 = 64, D1 = 256, D2 = 128, D3 = 64, DH = 16;   // head: 12 mean + 1 value + 3 pad

// packed parameter blob (global, produced by lm_mlp_pack; mirrored 1:1 in shared memory)
constexpr int W1_OFF = 0;
constexpr int W2_OFF = W1_OFF + operand_bytes(D1, D0);
constexpr int W3_OFF = W2_OFF + operand_bytes(D2, D1);
constexpr int WH_OFF = W3_OFF + operand_bytes(D3, D2);
constexpr int BIAS_OFF = WH_OFF + operand_bytes(DH, D3);          // fp32 biases: 256 + 128 + 64 + 16
constexpr int SBIAS_OFF = BIAS_OFF + (D1 + D2 + D3 + DH) * 4;      // log2(e) x the hidden-layer biases (inference kernel, see kScaled)
constexpr int BLOB_BYTES = SBIAS_OFF + (D1 + D2 + D3) * 4;
static_assert(BLOB_BYTES % 16 == 0 && W2_OFF % 16 == 0 && W3_OFF % 16 == 0 && WH_OFF % 16 == 0 && BIAS_OFF % 16 == 0 && SBIAS_OFF % 16 == 0, "align");

// activation regions (bf16 A operands).  A1 [128 x 256] = four 64-column K chunks (slots) that the layer-1 epilogue produces one by
// one and layer 2 consumes at once.  Slots 2-3 are ALIASED onto A2: A2 is only live from E2(t) until layer 3 of tile t has read
// it, A1's chunks 2-3 from E1's sub-steps 2-3 until layer 2 has read them (= before E2 starts), so the two never overlap; the only
// condition — layer 3 of tile t-1 complete before sub-step 2 of E1(t) stores — holds long before that store is reached.  (A two-slot
// ring in the same memory made sub-step s wait for the layer-2 MMAs of chunk s-2: 20.3 -> 18.4 us per launch at 65,536 envs.)
constexpr int CHUNK_BYTES = operand_bytes(TM, 64);                // one 64-wide K chunk of an A operand (= A0, A3)
constexpr int A1_OFF = BLOB_BYTES;                                // A1 slots 0..3
constexpr int A2_OFF = A1_OFF + 2 * CHUNK_BYTES;                  // A2 [128 x 128] = slots 2-3 of A1
constexpr int A0_OFF = A2_OFF + operand_bytes(TM, D2);            // A0 [128 x 64]
constexpr int A3_OFF = A0_OFF + CHUNK_BYTES;                      // A3 [128 x 64]
constexpr int SMEM_BYTES = A3_OFF + CHUNK_BYTES + 64;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");
static_assert(A1_OFF % 16 == 0 && CHUNK_BYTES % 16 == 0, "align");
// TMEM accumulator columns
constexpr uint32_t ACC1_COL = 0, ACC2_COL = 256, ACC3_COL = 384, ACCH_COL = 448, TMEM_COLS = 512;

// four K/16 MMAs of one layer (single thread): k-steps [ks0, ks0 + 4) of the B operand; the A operand's K chunk starts at
// `a_saddr` (k-step 0 there)
template <int kN>
__device__ __forceinline__ void issue_ksteps(uint32_t tmem_d, uint32_t a_saddr, uint32_t b_saddr, int ks0) {
  constexpr uint32_t idesc = instr_desc(TM, kN);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint64_t ad = smem_desc(a_saddr + i * 2 * lbo_bytes(TM), lbo_bytes(TM));
    const uint64_t bd = smem_desc(b_saddr + (ks0 + i) * 2 * lbo_bytes(kN), lbo_bytes(kN));
    umma_bf16(tmem_d, ad, bd, idesc, (ks0 + i) > 0 ? 1u : 0u);
  }
}

// 16 accumulator columns [c0, c0+16) -> +bias, ELU, bf16 -> eight packed words (two 16-byte K chunks of the next A operand).
// kScaled (the inference kernel): activations are stored multiplied by log2(e) and `bias` holds log2(e) b (lm_tc.cuh, elu_pair_scaled).
template <bool kScaled>
__device__ __forceinline__ void elu16(const uint32_t* acc, const float* __restrict__ bias, int c0, uint32_t (&w)[8]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 b = *reinterpret_cast<const float4*>(bias + c0 + 4 * k);
    if (kScaled) {
      w[2 * k] = elu_pair_mix_scaled<LM_MLP_FMA_PAIRS>((2 * k) & 3, acc[4 * k], acc[4 * k + 1], b.x, b.y);
      w[2 * k + 1] = elu_pair_mix_scaled<LM_MLP_FMA_PAIRS>((2 * k + 1) & 3, acc[4 * k + 2], acc[4 * k + 3], b.z, b.w);
    } else {
      // LM_MLP_FMA_PAIRS of every 4 ELU pairs take their exponential on the FMA pipe instead of MUFU (lm_tc.cuh)
      w[2 * k] = elu_pair_mix<LM_MLP_FMA_PAIRS>((2 * k) & 3, acc[4 * k], acc[4 * k + 1], b.x, b.y);
      w[2 * k + 1] = elu_pair_mix<LM_MLP_FMA_PAIRS>((2 * k + 1) & 3, acc[4 * k + 2], acc[4 * k + 3], b.z, b.w);
    }
  }
}
// `gdst` (training forward only): the same two chunks into the tile's activation blob in global memory (same layout, same offsets)
__device__ __forceinline__ void store16(const uint32_t (&w)[8], uint8_t* dst, int row, int c0, uint8_t* gdst = nullptr) {
  *reinterpret_cast<uint4*>(dst + elem_offset(TM, row, c0)) = make_uint4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<uint4*>(dst + elem_offset(TM, row, c0 + 8)) = make_uint4(w[4], w[5], w[6], w[7]);
  if (gdst != nullptr) {
    __stcs(reinterpret_cast<uint4*>(gdst + elem_offset(TM, row, c0)), make_uint4(w[0], w[1], w[2], w[3]));
    __stcs(reinterpret_cast<uint4*>(gdst + elem_offset(TM, row, c0 + 8)), make_uint4(w[4], w[5], w[6], w[7]));
  }
}

// Hidden-layer epilogue of one warp.  The layer's accumulator is handed to the next layer in kMacro steps; in step m the 16 warps
// together cover accumulator columns [W m, W (m + 1)), W = 64 kPieces (this warp: 16 kPieces of them, as kPieces TMEM loads of 16
// columns), i.e. kPieces 64-wide K chunks of the next layer's A operand, then publish them with ONE proxy fence and ONE arrival on
// `bars[m]` (LM_MLP_PIECES = 1: a fence per 16 columns, 7 per tile; 2: 4 per tile).  The ELU arithmetic of the next piece is
// written ahead of the fence, with the TMEM load after it issued as soon as its registers are free.
// `hook(k)` runs once per piece k before the following fence (shared-memory stores made there are published by the same fence).
// `alias_bar`: waited before the first store into A1 slot 2 (= A2's memory, see above).
struct NoHook { __device__ __forceinline__ void operator()(int) const {} };
template <int kMacro, int kPieces, bool kScaled, typename Hook = NoHook>
__device__ __forceinline__ void epilogue_hidden(uint32_t tmem_acc, const float* __restrict__ bias, uint8_t* dst, uint64_t* bars, int g, int q,
                                                int lane, uint8_t* gsave = nullptr, uint64_t* alias_bar = nullptr, uint32_t alias_parity = 0,
                                                Hook hook = Hook()) {
  constexpr int kN = kMacro * kPieces;
  const int row = g * 32 + lane;
  // piece k = (m, p): columns c0(k) = 64 kPieces m + 16 kPieces q + 16 p
  auto c0 = [&](int k) { return 64 * kPieces * (k / kPieces) + 16 * kPieces * q + 16 * (k % kPieces); };
  const uint32_t tbase = tmem_acc + ((uint32_t)(g * 32) << 16);
  uint32_t a[16], b[16], w[8];
  tmem_ld16(tbase + c0(0), a);
  tmem_wait();
  if (kN > 1) tmem_ld16(tbase + c0(1), b);
  elu16<kScaled>(a, bias, c0(0), w);
#pragma unroll
  for (int k = 0; k < kN; ++k) {
    uint32_t* nxt = (k & 1) ? a : b;          // holds piece k + 1 (in flight or landed)
    uint32_t* free_buf = (k & 1) ? b : a;     // piece k: already consumed into w
    const int c = c0(k);
    if (alias_bar != nullptr && c / 64 >= 2 && (k == 0 || c0(k - 1) / 64 < 2)) mbar_wait(alias_bar, alias_parity);
    store16(w, dst + (c / 64) * CHUNK_BYTES, row, c % 64, gsave != nullptr ? gsave + (c / 64) * CHUNK_BYTES : nullptr);
    if (k + 1 < kN) {
      tmem_wait();
#ifdef LM_DIAG_HALF_LDTM      // diagnostic (wrong results): every other TMEM load skipped -> is the kernel TMEM-read bound?
      if (k + 2 < kN && ((k + 2) & 1) == 0) tmem_ld16(tbase + c0(k + 2), free_buf);
#else
      if (k + 2 < kN) tmem_ld16(tbase + c0(k + 2), free_buf);
#endif
      elu16<kScaled>(nxt, bias, c0(k + 1), w);
    }
    hook(k);
    if (k % kPieces == kPieces - 1) {
      fence_proxy_async();                     // generic-proxy A operand writes -> visible to the tensor core's async proxy
      tc_fence_before();                       // orders this warp's tcgen05.ld before the MMA that reuses the accumulator
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[k / kPieces]);
    }
  }
}

struct Args {
  const uint8_t* blob;
  const float* obs;
  const float* norm_mean;      // may be NULL
  const float* norm_inv_std;   // may be NULL
  float norm_clip;
  float* mu;
  float* value;
  int n;
  int num_tiles;
  long long* timeline;         // diagnostic: clock64 stamps of CTA 0 (NULL in production launches)
  // training forward (k_mlp_forward<true>): the bf16 operands of every layer saved per tile for the backward kernels (lm_mlp_train.cu)
  uint8_t *save_a0, *save_a1, *save_a2, *save_a3;
  // evaluation with an action sink: the head epilogue also delivers action = mean into the task's buffers (lm_sink.cuh)
  LmActionSink sink; int has_sink;
};

// diagnostic timeline (lm_mlp_set_timeline): slot k <- clock64() of CTA 0.  Compiled into the kStamp instantiation only, which is
// launched while a timeline buffer is registered: the production kernel carries none of these instructions.
template <bool kStamp>
__device__ __forceinline__ void stamp(const Args& a, int k) {
  if (kStamp && a.timeline != nullptr && blockIdx.x == 0) a.timeline[k] = clock64();
}

struct ObsRegs { float4 v[4]; };

// thread tid of the 512 epilogue threads owns float4 (j*512 + tid) of the [128 x 64] fp32 tile: row j*32 + tid/16, columns 4*(tid%16)..
__device__ __forceinline__ void obs_fetch(const Args& a, int tile, int tid, ObsRegs& r) {
  const int rows = min(TM, a.n - tile * TM);
  const float* base = a.obs + (size_t)tile * TM * D0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = j * 32 + (tid >> 4);
    r.v[j] = row < rows ? ld_nc_f4(base + (size_t)(j * E_THREADS + tid) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// One float4 of the thread's share: normalise, bf16, clamp, into the A0 operand (and the training forward's saved copy).  The clamp is
// taken on the packed bf16 pairs — rounding is monotone, so min(bf16(x), bf16(c)) = bf16(min(x, c)) for any clip c — and `dst` is the
// thread's precomputed A0 address of the chunk (base + j * 512 bytes), so a piece costs 15 instructions.  kGuard (training forward,
// whose saved rows feed the weight gradients): rows beyond the batch stay zero.
template <bool kGuard, bool kScaled>
__device__ __forceinline__ void obs_convert_one(float4 x, bool row_valid, const float4& mean, const float4& inv_std, uint32_t clip2, uint32_t nclip2,
                                                bool has_norm, uint8_t* dst, uint8_t* gdst) {
  // kScaled: the caller folded log2(e) into inv_std and the clip bounds; without normalisation the scaling is applied here
  if (has_norm) {
    x.x = (x.x - mean.x) * inv_std.x; x.y = (x.y - mean.y) * inv_std.y;
    x.z = (x.z - mean.z) * inv_std.z; x.w = (x.w - mean.w) * inv_std.w;
  } else if (kScaled) {
    x.x *= 1.4426950408889634f; x.y *= 1.4426950408889634f; x.z *= 1.4426950408889634f; x.w *= 1.4426950408889634f;
  }
  uint32_t lo = pack_bf16(x.x, x.y), hi = pack_bf16(x.z, x.w);
  if (has_norm) {
    asm("min.bf16x2 %0, %0, %1;" : "+r"(lo) : "r"(clip2));
    asm("max.bf16x2 %0, %0, %1;" : "+r"(lo) : "r"(nclip2));
    asm("min.bf16x2 %0, %0, %1;" : "+r"(hi) : "r"(clip2));
    asm("max.bf16x2 %0, %0, %1;" : "+r"(hi) : "r"(nclip2));
  }
  if (kGuard && !row_valid) { lo = 0u; hi = 0u; }
  *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
  if (gdst != nullptr) *reinterpret_cast<uint2*>(gdst) = make_uint2(lo, hi);
}

// Head epilogue, spread over the four column-quarter warps of a lane group: quarter q reads accumulator columns [4q, 4q + 4) —
// q = 0..2: four policy means -> mu (and, with an action sink, the same chunk into the task's action ring / joint targets),
// q = 3: column 12 -> value.  Every warp does a quarter of the work instead of one warp per scheduler doing all of it.
template <bool kScaled>
__device__ __forceinline__ void epilogue_head(const Args& a, uint32_t tmem_acc, const float* __restrict__ biash, int tile, int g, int lane, int q) {
  uint32_t acc[4];
  tmem_ld4(tmem_acc + ((uint32_t)(g * 32) << 16) + (uint32_t)(4 * q), acc);
  tmem_wait();
  tc_fence_before();
  const int row0 = tile * TM, rows = min(TM, a.n - row0);
  const int row = g * 32 + lane;
  if (row < rows) {
    const float4 b = *reinterpret_cast<const float4*>(biash + 4 * q);
    // kScaled: the accumulator holds log2(e) x (W h): one FMA per output undoes the scaling and adds the bias
    constexpr float kInv = kScaled ? 0.6931471805599453f : 1.0f;
    const float4 o = kScaled ? make_float4(fmaf(__uint_as_float(acc[0]), kInv, b.x), fmaf(__uint_as_float(acc[1]), kInv, b.y),
                                           fmaf(__uint_as_float(acc[2]), kInv, b.z), fmaf(__uint_as_float(acc[3]), kInv, b.w))
                             : make_float4(__uint_as_float(acc[0]) + b.x, __uint_as_float(acc[1]) + b.y, __uint_as_float(acc[2]) + b.z,
                                           __uint_as_float(acc[3]) + b.w);
    if (q < 3) {
      reinterpret_cast<float4*>(a.mu + (size_t)(row0 + row) * 12)[q] = o;
      if (a.has_sink) sink_store(a.sink, row0 + row, q, o);
    } else {
      a.value[row0 + row] = o.x;
    }
  }
}

// The inference instantiation is compiled under a 672-thread bound, i.e. for at most 80 registers (it is launched with 544 threads):
// registers are allocated in units of four warps, so the 17-warp CTA then leaves room for one 128-thread CTA of another kernel on
// the SM — the reset half of the next pre_physics_step, which runs beside this kernel (lm_task_set_early_reset).  At 96 registers
// the kernel alone is 0.7 us faster (17.4 vs 18.1 us at 65,536 envs) but nothing fits beside it: 42.9 vs 37.6 us per env step.
#ifndef LM_MLP_LAUNCH_BOUND
#define LM_MLP_LAUNCH_BOUND 672
#endif
template <bool kSave, bool kStamp = false>
__global__ void __launch_bounds__(kSave ? THREADS : LM_MLP_LAUNCH_BOUND, 1) k_mlp_forward(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // mbarriers.  count 1: bar_w (weight TMA), bar_acc* (tcgen05.commit: accumulator complete).  count 16 (one arrive per epilogue
  // warp): bar_a0 (A0 operand written), bar_act1[s] / bar_act2[s] / bar_act3 (K chunk s of the next layer's A operand written; for
  // the last sub-step also: accumulator fully read).  Every barrier completes exactly once per tile, so the parity of a wait is the
  // parity of the CTA-local tile index.
  __shared__ uint64_t bar_w, bar_a0, bar_acc1, bar_acc2, bar_acc3, bar_acch, bar_act1[4 / kPc], bar_act2[2 / kPc], bar_act3[1];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* A1 = smem + A1_OFF;
  uint8_t* A2 = smem + A2_OFF;
  uint8_t* A0 = smem + A0_OFF;
  uint8_t* A3 = smem + A3_OFF;
  // Inference (kScaled): hidden activations are stored multiplied by log2(e), which removes the multiply in front of every ex2
  // (elu_pair_scaled); the training forward keeps the plain form, because its saved activations feed the backward kernels.
  constexpr bool kScaled = !kSave;
  const float* bias1 = reinterpret_cast<const float*>(smem + (kScaled ? SBIAS_OFF : BIAS_OFF));
  const float* bias2 = bias1 + D1;
  const float* bias3 = bias2 + D2;
  const float* biash = reinterpret_cast<const float*>(smem + BIAS_OFF) + D1 + D2 + D3;

  if (tid == 0) {
    mbar_init(&bar_w, 1); mbar_init(&bar_a0, E_WARPS);
    mbar_init(&bar_acc1, 1); mbar_init(&bar_acc2, 1); mbar_init(&bar_acc3, 1); mbar_init(&bar_acch, 1);
    for (int s = 0; s < 4 / kPc; ++s) mbar_init(&bar_act1[s], E_WARPS);
    for (int s = 0; s < 2 / kPc; ++s) mbar_init(&bar_act2[s], E_WARPS);
    mbar_init(&bar_act3[0], E_WARPS);
    fence_mbar_init();
  }
  if (warp == E_WARPS) tmem_alloc(&tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) stamp<kStamp>(a, 0);
  const int tile0 = blockIdx.x, stride = gridDim.x;
  const int n_loc = tile0 < a.num_tiles ? (a.num_tiles - tile0 + stride - 1) / stride : 0;   // tiles of this CTA

  // Software pipeline over the CTA's tiles (iteration `it` = tile t).  Epilogue warps run, per iteration,
  //   E1(t) [+ obs(t+1) -> A0, one float4 per sub-step] -> heads(t-2) -> E3(t-1) -> E2(t)
  // so that every wait on a tensor-core result (acc2(t) after E1(t), acc3(t) after E2(t), heads(t) after E3(t)) has a
  // full layer's worth of other work in front of it.  The control thread issues MMAs in the matching order.
  if (warp == E_WARPS) {
    // ================================================================== control warp: one thread issues TMA + every MMA
    if (elect_one()) {
      mbar_expect_tx(&bar_w, BLOB_BYTES);
      tma_load_1d(smem, a.blob, BLOB_BYTES, &bar_w);      // resident weights + biases: one bulk copy per CTA
      mbar_wait(&bar_w, 0);
      stamp<kStamp>(a, 1);
      const uint32_t sW1 = smem_u32(smem + W1_OFF), sW2 = smem_u32(smem + W2_OFF), sW3 = smem_u32(smem + W3_OFF), sWH = smem_u32(smem + WH_OFF);
      const uint32_t sA1 = smem_u32(A1), sA2 = smem_u32(A2), sA0 = smem_u32(A0), sA3 = smem_u32(A3);
      if (n_loc > 0) {
        mbar_wait(&bar_a0, 0);
        tc_fence_after();
        issue_ksteps<D1>(tmem + ACC1_COL, sA0, sW1, 0);
        umma_commit(&bar_acc1);
      }
      for (int it = 0; it < n_loc; ++it) {
        const uint32_t ph = it & 1;
        // layer 2 of tile t: K chunk s as soon as the layer-1 epilogue has produced it
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (s % kPc == 0) {
            mbar_wait(&bar_act1[s / kPc], ph);
            tc_fence_after();
          }
          issue_ksteps<D2>(tmem + ACC2_COL, sA1 + s * CHUNK_BYTES, sW2, 4 * s);
        }
        umma_commit(&bar_acc2);
        // layer 1 of tile t+1 (its accumulator columns were released by the last layer-1 sub-step above)
        if (it + 1 < n_loc) {
          mbar_wait(&bar_a0, ph ^ 1);
          tc_fence_after();
          issue_ksteps<D1>(tmem + ACC1_COL, sA0, sW1, 0);
          umma_commit(&bar_acc1);
        }
        // heads of tile t-1
        if (it >= 1) {
          mbar_wait(&bar_act3[0], ph ^ 1);
          tc_fence_after();
          issue_ksteps<DH>(tmem + ACCH_COL, sA3, sWH, 0);
          umma_commit(&bar_acch);
        }
        // layer 3 of tile t
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (s % kPc == 0) {
            mbar_wait(&bar_act2[s / kPc], ph);
            tc_fence_after();
          }
          issue_ksteps<D3>(tmem + ACC3_COL, sA2 + s * CHUNK_BYTES, sW3, 4 * s);
        }
        umma_commit(&bar_acc3);
      }
      if (n_loc > 0) {                                    // heads of the last tile
        mbar_wait(&bar_act3[0], (n_loc - 1) & 1);
        tc_fence_after();
        issue_ksteps<DH>(tmem + ACCH_COL, sA3, sWH, 0);
        umma_commit(&bar_acch);
      }
    }
  } else {
    // ================================================================== epilogue warps
    const int g = warp & 3, q = warp >> 2;
    const bool has_norm = a.norm_mean != nullptr;
    // training forward: tile t's saved operand of layer L lives at save_aL + t * operand_bytes(128, D_L)
    auto g0 = [&](int t) -> uint8_t* { return kSave ? a.save_a0 + (size_t)t * operand_bytes(TM, D0) : nullptr; };
    auto g1 = [&](int t) -> uint8_t* { return kSave ? a.save_a1 + (size_t)t * operand_bytes(TM, D1) : nullptr; };
    auto g2 = [&](int t) -> uint8_t* { return kSave ? a.save_a2 + (size_t)t * operand_bytes(TM, D2) : nullptr; };
    auto g3 = [&](int t) -> uint8_t* { return kSave ? a.save_a3 + (size_t)t * operand_bytes(TM, D3) : nullptr; };
    float4 mean = make_float4(0.f, 0.f, 0.f, 0.f), inv_std = make_float4(1.f, 1.f, 1.f, 1.f);
    if (has_norm) {
      mean = *reinterpret_cast<const float4*>(a.norm_mean + 4 * (tid & 15));
      inv_std = *reinterpret_cast<const float4*>(a.norm_inv_std + 4 * (tid & 15));
      if (kScaled) { inv_std.x *= 1.4426950408889634f; inv_std.y *= 1.4426950408889634f; inv_std.z *= 1.4426950408889634f; inv_std.w *= 1.4426950408889634f; }
    }
    ObsRegs regs;
    // this thread's chunk of A0: row (tid / 16) + 32 j, columns 4 (tid % 16)..: byte offset of j = 0, then 512 bytes per j
    const int a0_off = elem_offset(TM, tid >> 4, 4 * (tid & 15));
    const float osc = kScaled ? 1.4426950408889634f : 1.0f;
    const uint32_t clip2 = pack_bf16(a.norm_clip * osc, a.norm_clip * osc), nclip2 = pack_bf16(-a.norm_clip * osc, -a.norm_clip * osc);
    pdl_wait();                                            // the observations are the predecessor kernel's output
    // Dependents are released only now, i.e. once the producer of the observations (the post step) has completed: a dependent that
    // needs nothing from THIS kernel (lm_task_set_early_reset: the reset half of the next pre step) may then run beside it.
    pdl_launch_dependents();
    if (n_loc > 0) {
      obs_fetch(a, tile0, tid, regs);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        obs_convert_one<kSave, kScaled>(regs.v[j], j * 32 + (tid >> 4) < a.n - tile0 * TM, mean, inv_std, clip2, nclip2, has_norm, A0 + a0_off + j * 512,
                               kSave ? g0(tile0) + a0_off + j * 512 : nullptr);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_a0);
      if (n_loc > 1) obs_fetch(a, tile0 + stride, tid, regs);
    }
    const bool st = tid == 0;
    if (st) stamp<kStamp>(a, 2);
    mbar_wait(&bar_w, 0);                                  // biases live in the weight blob
    for (int it = 0; it < n_loc; ++it) {
      const uint32_t ph = it & 1;
      const int tile = tile0 + it * stride;
      const bool more = it + 1 < n_loc;
      // ---- E1(t): TMEM[128 x 256] -> A1, four K chunks.  Layer 1 of tile t has consumed A0 (bar_acc1), so sub-step s also converts
      // float4 s of this thread's share of tile t+1's observations into A0: the conversion rides in the MUFU-bound sub-steps' spare
      // issue slots and its stores are published by the epilogue's own fences.
      if (st) stamp<kStamp>(a, 8 + 8 * it + 0);
      mbar_wait(&bar_acc1, ph);
      tc_fence_after();
      if (st) stamp<kStamp>(a, 8 + 8 * it + 1);
      uint8_t* gA0n = more ? g0(tile + stride) : nullptr;
      const int rows_n = min(TM, a.n - (tile + stride) * TM);
      auto obs_hook = [&](int s) {
        if (more)
          obs_convert_one<kSave, kScaled>(regs.v[s], s * 32 + (tid >> 4) < rows_n, mean, inv_std, clip2, nclip2, has_norm, A0 + a0_off + s * 512,
                                 kSave ? gA0n + a0_off + s * 512 : nullptr);
      };
      epilogue_hidden<4 / kPc, kPc, kScaled>(tmem + ACC1_COL, bias1, A1, bar_act1, g, q, lane, g1(tile), it >= 1 ? &bar_acc3 : nullptr, ph ^ 1, obs_hook);
      if (st) stamp<kStamp>(a, 8 + 8 * it + 2);
      if (more) {
        if (lane == 0) mbar_arrive(&bar_a0);               // behind the last fence of E1
        if (it + 2 < n_loc) obs_fetch(a, tile + 2 * stride, tid, regs);
      }
      if (st) stamp<kStamp>(a, 8 + 8 * it + 3);
      // ---- heads of tile t-2 (every warp waits: A3 must be drained before E3(t-1) overwrites it)
      if (it >= 2) {
        mbar_wait(&bar_acch, ph);
        tc_fence_after();
        epilogue_head<kScaled>(a, tmem + ACCH_COL, biash, tile - 2 * stride, g, lane, q);
      }
      // ---- E3(t-1): TMEM[128 x 64] -> A3
      if (st) stamp<kStamp>(a, 8 + 8 * it + 4);
      if (it >= 1) {
        mbar_wait(&bar_acc3, ph ^ 1);
        tc_fence_after();
        epilogue_hidden<1, 1, kScaled>(tmem + ACC3_COL, bias3, A3, bar_act3, g, q, lane, g3(tile - stride));
      }
      // ---- E2(t): TMEM[128 x 128] -> A2 (layer 3 of tile t-1 has finished reading it: bar_acc3 above)
      if (st) stamp<kStamp>(a, 8 + 8 * it + 5);
      mbar_wait(&bar_acc2, ph);
      tc_fence_after();
      if (st) stamp<kStamp>(a, 8 + 8 * it + 6);
      epilogue_hidden<2 / kPc, kPc, kScaled>(tmem + ACC2_COL, bias2, A2, bar_act2, g, q, lane, g2(tile));
      if (st) stamp<kStamp>(a, 8 + 8 * it + 7);
    }
    // ---- drain: heads(n-2), E3(n-1), heads(n-1)
    if (n_loc >= 2) {
      mbar_wait(&bar_acch, n_loc & 1);
      tc_fence_after();
      epilogue_head<kScaled>(a, tmem + ACCH_COL, biash, tile0 + (n_loc - 2) * stride, g, lane, q);
    }
    if (n_loc >= 1) {
      mbar_wait(&bar_acc3, (n_loc - 1) & 1);
      tc_fence_after();
      epilogue_hidden<1, 1, kScaled>(tmem + ACC3_COL, bias3, A3, bar_act3, g, q, lane, g3(tile0 + (n_loc - 1) * stride));
      mbar_wait(&bar_acch, (n_loc - 1) & 1);
      tc_fence_after();
      epilogue_head<kScaled>(a, tmem + ACCH_COL, biash, tile0 + (n_loc - 1) * stride, g, lane, q);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) stamp<kStamp>(a, 3);
  if (warp == E_WARPS) tmem_dealloc(tmem, TMEM_COLS);
}

// fp32 nn.Linear parameters -> the resident blob (bf16 canonical operands + fp32 biases) in ONE launch: the PPO update re-packs
// after every optimiser step (the fp32 masters are written by lm_adam_step, the kernels read this blob).
struct PackArgs {
  const float *w1, *b1, *w2, *b2, *w3, *b3, *w_mu, *b_mu, *w_value, *b_value;
  uint8_t* blob;
};
constexpr int PK_W1 = D1 * D0, PK_W2 = PK_W1 + D2 * D1, PK_W3 = PK_W2 + D3 * D2, PK_WH = PK_W3 + DH * D3, PK_ALL = PK_WH + D1 + D2 + D3 + DH;
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ void pack_elem(const float* __restrict__ w, uint8_t* dst, int rows_total, int k, int i) {
  const int row = i / k, col = i - row * k;
  *reinterpret_cast<__nv_bfloat16*>(dst + elem_offset(rows_total, row, col)) = __float2bfloat16_rn(w[i]);
}
__global__ void k_pack_all(const PackArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < PK_W1) pack_elem(a.w1, a.blob + W1_OFF, D1, D0, i);
  else if (i < PK_W2) pack_elem(a.w2, a.blob + W2_OFF, D2, D1, i - PK_W1);
  else if (i < PK_W3) pack_elem(a.w3, a.blob + W3_OFF, D3, D2, i - PK_W2);
  else if (i < PK_WH) {                       // head rows 0..11 = mean weights, row 12 = value weights, 13..15 zero
    const int j = i - PK_W3, row = j / D3, col = j - row * D3;
    const float v = row < 12 ? a.w_mu[j] : (row == 12 ? a.w_value[col] : 0.0f);
    *reinterpret_cast<__nv_bfloat16*>(a.blob + WH_OFF + elem_offset(DH, row, col)) = __float2bfloat16_rn(v);
  } else if (i < PK_ALL) {
    const int j = i - PK_WH;
    float* dst = reinterpret_cast<float*>(a.blob + BIAS_OFF);
    float* sdst = reinterpret_cast<float*>(a.blob + SBIAS_OFF);
    if (j < D1) { dst[j] = a.b1[j]; sdst[j] = a.b1[j] * kLog2e; }
    else if (j < D1 + D2) { dst[j] = a.b2[j - D1]; sdst[j] = a.b2[j - D1] * kLog2e; }
    else if (j < D1 + D2 + D3) { dst[j] = a.b3[j - D1 - D2]; sdst[j] = a.b3[j - D1 - D2] * kLog2e; }
    else if (j < D1 + D2 + D3 + 12) dst[j] = a.b_mu[j - D1 - D2 - D3];
    else dst[j] = j == D1 + D2 + D3 + 12 ? a.b_value[0] : 0.0f;
  }
}

}  // namespace mlp
}  // namespace lm

using namespace lm;
using namespace lm::mlp;

extern "C" {

LmMlp* lm_mlp_create(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { lm_set_error("lm_mlp_create: %s", cudaGetErrorString(e)); return nullptr; }
  if (prop.major != 10) { lm_set_error("lm_mlp_create: tcgen05 kernel needs sm_100a, device is sm_%d%d", prop.major, prop.minor); return nullptr; }
  cudaSetDevice(device);
  LmMlp* m = new LmMlp();
  m->device = device; m->sm_count = prop.multiProcessorCount; m->packed = false; m->blob = nullptr; m->blob_t = nullptr; m->timeline = nullptr; m->fresh_pack = true;
  if (cudaMalloc(&m->blob, BLOB_BYTES) != cudaSuccess) { lm_set_error("lm_mlp_create: cudaMalloc failed"); delete m; return nullptr; }
  cudaMemset(m->blob, 0, BLOB_BYTES);
  e = cudaFuncSetAttribute(k_mlp_forward<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_forward<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_forward<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) { lm_set_error("lm_mlp_create: smem attribute: %s", cudaGetErrorString(e)); cudaFree(m->blob); delete m; return nullptr; }
  return m;
}

void lm_mlp_destroy(LmMlp* m) {
  if (!m) return;
  cudaFree(m->blob);
  lm_mlp_train_release(m);
  delete m;
}

int lm_mlp_set_timeline(LmMlp* m, long long* stamps_dev) {
  if (!m) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_set_timeline: NULL");
  m->timeline = stamps_dev;
  return LM_OK;
}

int lm_mlp_pack(LmMlp* m, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3,
                const float* w_mu, const float* b_mu, const float* w_value, const float* b_value, void* stream) {
  if (!m || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !w_mu || !b_mu || !w_value || !b_value)
    return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_pack: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  // the blob's padding bytes (zeroed at creation) are never read and every element is rewritten: no memset between packs
  const PackArgs pa{w1, b1, w2, b2, w3, b3, w_mu, b_mu, w_value, b_value, m->blob};
  k_pack_all<<<(PK_ALL + 255) / 256, 256, 0, s>>>(pa);
  const int rc = lm_mlp_train_pack(m, w2, w3, w_mu, w_value, stream);      // transposed operands of the backward kernel
  if (rc != LM_OK) return rc;
  m->packed = true;
  m->fresh_pack = true;
  return lm_check_launch("lm_mlp_pack");
}

int lm_mlp_forward(LmMlp* m, const float* obs, const float* norm_mean, const float* norm_inv_std, float norm_clip, float* mu,
                   float* value, int32_t n, void* stream) {
  return lm_mlp_forward_sink(m, obs, norm_mean, norm_inv_std, norm_clip, mu, value, n, nullptr, stream);
}

int lm_mlp_forward_sink(LmMlp* m, const float* obs, const float* norm_mean, const float* norm_inv_std, float norm_clip, float* mu,
                        float* value, int32_t n, const LmActionSink* sink, void* stream) {
  if (!m || !obs || !mu || !value || n <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_forward: bad argument");
  if (sink && (!sink->ring_next || (((uintptr_t)sink->ring_next | (uintptr_t)sink->targets_loco | (uintptr_t)sink->targets_mani) & 15)))
    return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_forward_sink: sink buffers must be non-NULL / 16-byte aligned");
  if (!m->packed) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_forward: weights not packed (call lm_mlp_pack)");
  if ((((uintptr_t)obs) | ((uintptr_t)mu)) & 15) return lm_fail(LM_ERR_ALIGNMENT, "lm_mlp_forward: obs/mu must be 16-byte aligned");
  if ((norm_mean == nullptr) != (norm_inv_std == nullptr)) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_forward: give both norm vectors or neither");
  Args a;
  a.blob = m->blob; a.obs = obs; a.norm_mean = norm_mean; a.norm_inv_std = norm_inv_std; a.norm_clip = norm_clip;
  a.mu = mu; a.value = value; a.n = n; a.num_tiles = (n + TM - 1) / TM; a.timeline = m->timeline;
  a.save_a0 = a.save_a1 = a.save_a2 = a.save_a3 = nullptr;
  a.has_sink = sink != nullptr;
  if (sink) a.sink = *sink; else memset(&a.sink, 0, sizeof(a.sink));
  const int grid = a.num_tiles < m->sm_count ? a.num_tiles : m->sm_count;
  // The resident-weight TMA is issued before pdl_wait(): allowed only while the blob is not being rewritten upstream.
  const bool pdl = !m->fresh_pack;
  m->fresh_pack = false;
  cudaError_t le = a.timeline != nullptr
                       ? lm_launch(k_mlp_forward<false, true>, dim3(grid), dim3(THREADS), SMEM_BYTES, (cudaStream_t)stream, pdl, a)
                       : lm_launch(k_mlp_forward<false>, dim3(grid), dim3(THREADS), SMEM_BYTES, (cudaStream_t)stream, pdl, a);
  if (le != cudaSuccess) return lm_fail((int)le, "lm_mlp_forward: %s", cudaGetErrorString(le));
  return lm_check_launch("lm_mlp_forward");
}

// The same forward for the PPO update: additionally leaves the bf16 operands of every layer (A0 = normalised observations,
// A1, A2, A3) as tile blobs (lm_mlp_blob_bytes(n, D) bytes each) for lm_mlp_backward / lm_mlp_wgrad.
int lm_mlp_forward_train(LmMlp* m, const float* obs, const float* norm_mean, const float* norm_inv_std, float norm_clip, float* mu,
                         float* value, void* a0, void* a1, void* a2, void* a3, int32_t n, void* stream) {
  if (!m || !obs || !mu || !value || !a0 || !a1 || !a2 || !a3 || n <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_forward_train: bad argument");
  if (!m->packed) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_forward_train: weights not packed (call lm_mlp_pack)");
  if ((((uintptr_t)obs) | ((uintptr_t)mu) | ((uintptr_t)a0) | ((uintptr_t)a1) | ((uintptr_t)a2) | ((uintptr_t)a3)) & 15)
    return lm_fail(LM_ERR_ALIGNMENT, "lm_mlp_forward_train: obs / mu / activation blobs must be 16-byte aligned");
  if ((norm_mean == nullptr) != (norm_inv_std == nullptr)) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_forward_train: give both norm vectors or neither");
  Args a;
  a.blob = m->blob; a.obs = obs; a.norm_mean = norm_mean; a.norm_inv_std = norm_inv_std; a.norm_clip = norm_clip;
  a.mu = mu; a.value = value; a.n = n; a.num_tiles = (n + TM - 1) / TM; a.timeline = nullptr;
  a.save_a0 = (uint8_t*)a0; a.save_a1 = (uint8_t*)a1; a.save_a2 = (uint8_t*)a2; a.save_a3 = (uint8_t*)a3;
  a.has_sink = 0; memset(&a.sink, 0, sizeof(a.sink));
  const int grid = a.num_tiles < m->sm_count ? a.num_tiles : m->sm_count;
  m->fresh_pack = false;
  cudaError_t le = lm_launch(k_mlp_forward<true>, dim3(grid), dim3(THREADS), SMEM_BYTES, (cudaStream_t)stream, false, a);
  if (le != cudaSuccess) return lm_fail((int)le, "lm_mlp_forward_train: %s", cudaGetErrorString(le));
  return lm_check_launch("lm_mlp_forward_train");
}

}  // extern "C"
