// PPO update of the MLP actor-critic on tcgen05: weight-gradient GEMMs over saved activation / gradient tiles.
//
// Tile format ("blob"): a [B, D] bf16 matrix is stored tile by tile (128 samples), each tile in the UMMA canonical K-major
// no-swizzle operand layout of lm_tc.cuh, element (row, k) at (k/8) * 2064 + row * 16 + (k%8) * 2.  The same bytes are the MN-major
// INTERLEAVE canonical layout ((8,1,m),(8,k)) : ((1,8,SBO),(8,LBO)) of the transposed matrix with SBO = 2064 and LBO = 128
// (cute/atom/mma_traits_sm100.hpp), so a tile written once as the A operand of dA = dZ W (M = samples) is read here as the A
// (or B) operand of dW = dZ^T A (M = features, K = samples): no transpose pass exists anywhere.
//
// k_mlp_wgrad: persistent, one CTA per SM.  One elected thread streams the tiles' blobs into a two-slot shared-memory ring with
// TMA bulk copies and issues every MMA; all weight / bias gradients of the CTA's tiles accumulate in TMEM (504 of 512 columns)
// and are read out once at the end into per-CTA partials; k_wgrad_reduce sums the partials in a fixed order.
//   dW1 [256 x 64]  = dZ1^T A0   (two M = 128 halves)      db1 = dZ1^T 1   (ones operand, N = 8)
//   dW2 [128 x 256] = dZ2^T A1                            db2 = dZ2^T 1
//   dW3^T [128 x 64] = A2^T dZ3  (M = in: out = 64 < 128)  db3 = dZ3^T 1   (M = 128 reads 64 rows past the blob: ignored rows)
//   dWh^T [64 x 16] = A3^T dH    (rows 64.. ignored)       dbh = dH^T 1
#include <cuda_bf16.h>

#include "lm_comm.cuh"
#include "lm_common.cuh"
#include "lm_host.h"
#include "lm_mlp.h"
#include "lm_tc.cuh"

namespace lm {
namespace mlp_train {
using namespace tc;

constexpr int TM = 128;
constexpr int D0 = 64, D1 = 256, D2 = 128, D3 = 64, DH = 16;
__host__ __device__ constexpr int blob_bytes(int d) { return operand_bytes(TM, d); }

// ---------------------------------------------------------------- blob pack / unpack (tests, debugging, host-side glue)
__global__ void k_blob_pack(const float* __restrict__ x, uint8_t* __restrict__ blob, long long rows, int d) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long tiles = (rows + TM - 1) / TM;
  if (i >= tiles * TM * d) return;
  const long long row = i / d;
  const int k = (int)(i - row * d);
  const float v = row < rows ? x[row * d + k] : 0.0f;
  *reinterpret_cast<__nv_bfloat16*>(blob + (row / TM) * blob_bytes(d) + elem_offset(TM, (int)(row % TM), k)) = __float2bfloat16_rn(v);
}
__global__ void k_blob_unpack(const uint8_t* __restrict__ blob, float* __restrict__ x, long long rows, int d) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * d) return;
  const long long row = i / d;
  const int k = (int)(i - row * d);
  x[i] = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(blob + (row / TM) * blob_bytes(d) + elem_offset(TM, (int)(row % TM), k)));
}


// ---------------------------------------------------------------- backward through the activations
// k_mlp_backward: dH -> dZ3 -> dZ2 -> dZ1 per 128-sample tile, transposed weights resident in shared memory.
//   dA3 = dH Wh      (K = 16),  dZ3 = dA3 * ELU'(A3)
//   dA2 = dZ3 W3     (K = 64),  dZ2 = dA2 * ELU'(A2)
//   dA1 = dZ2 W2     (K = 128), dZ1 = dA1 * ELU'(A1)          ELU'(z) = 1 for a > 0, a + 1 otherwise (a = ELU(z), the saved operand)
// 16 epilogue warps (4 TMEM lane groups x 4 column quarters) + 1 control warp, as the forward kernel.  The saved activations are
// read straight from their global blobs (each element is used once, by the thread that owns it); every dZ leaves as a tile
// blob for lm_mlp_wgrad, dZ3 / dZ2 also stay in shared memory as the next GEMM's A operand.
constexpr int B_WARPS = 16, B_THREADS = B_WARPS * 32 + 32;
constexpr int WHT_OFF = 0;                                             // B operand of dA3: [N = 64 (in)][K = 16 (o)]
constexpr int W3T_OFF = WHT_OFF + operand_bytes(D3, DH);               // B operand of dA2: [N = 128][K = 64]
constexpr int W2T_OFF = W3T_OFF + operand_bytes(D2, D3);               // B operand of dA1: [N = 256][K = 128]
constexpr int WT_BYTES = W2T_OFF + operand_bytes(D1, D2);
constexpr int DH_OFF = WT_BYTES;
constexpr int DZ3_OFF = DH_OFF + blob_bytes(DH);
constexpr int DZ2_OFF = DZ3_OFF + blob_bytes(D3);
constexpr int BWD_SMEM_BYTES = DZ2_OFF + blob_bytes(D2) + 64;
static_assert(WT_BYTES % 16 == 0 && W3T_OFF % 16 == 0 && W2T_OFF % 16 == 0, "align");
static_assert(BWD_SMEM_BYTES <= 232448 - 1024, "exceeds 227 KB");
constexpr uint32_t BA3_COL = 0, BA2_COL = 64, BA1_COL = 192;           // 64 + 128 + 256 accumulator columns

struct BwdArgs {
  const uint8_t* wt;                        // transposed weight blob (lm_mlp_train_pack)
  const float* grad_mu; const float* grad_value;       // [n,12], [n]
  const uint8_t *a1, *a2, *a3;              // saved activations (lm_mlp_forward_train)
  uint8_t *dh, *dz3, *dz2, *dz1;            // out: tile blobs
  int n, num_tiles;
};

template <int kN, int kK>
__device__ __forceinline__ void issue_gemm(uint32_t acc, uint32_t a_saddr, uint32_t b_saddr) {      // D[128 x kN] = A[128 x kK] B^T, K-major operands
  constexpr uint32_t idesc = instr_desc(TM, kN);
#pragma unroll
  for (int ks = 0; ks < kK / 16; ++ks)
    umma_bf16(acc, smem_desc(a_saddr + ks * 2 * lbo_bytes(TM), lbo_bytes(TM)), smem_desc(b_saddr + ks * 2 * lbo_bytes(kN), lbo_bytes(kN)), idesc,
              (uint32_t)(ks > 0));
}

// 8 accumulator columns x the saved activation chunk -> dZ chunk (8 bf16)
__device__ __forceinline__ uint4 dz_chunk(const uint32_t* acc, const uint4& act) {
  const uint32_t aw[4] = {act.x, act.y, act.z, act.w};
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 a2 = *reinterpret_cast<const __nv_bfloat162*>(&aw[k]);
    const float a_lo = __low2float(a2), a_hi = __high2float(a2);
    const float d_lo = a_lo > 0.0f ? 1.0f : a_lo + 1.0f, d_hi = a_hi > 0.0f ? 1.0f : a_hi + 1.0f;
    w[k] = pack_bf16(__uint_as_float(acc[2 * k]) * d_lo, __uint_as_float(acc[2 * k + 1]) * d_hi);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// One backward epilogue of a warp: kCols accumulator columns starting at column q * kCols of `tmem_acc` -> dZ chunks to `sdst`
// (shared-memory operand, may be NULL) and `gdst` (global blob).  The saved activation chunks of the first kPre columns are fetched by
// act_prefetch BEFORE the wait on the accumulator, so their global latency hides behind the GEMM; the rest is fetched in the loop.
template <int kPre>
struct ActRegs { uint4 v[kPre / 8]; };
template <int kCols, int kPre>
__device__ __forceinline__ void act_prefetch(const uint8_t* __restrict__ gact, int g, int q, int lane, ActRegs<kPre>& r) {
  const int row = g * 32 + lane;
#pragma unroll
  for (int c = 0; c < kPre; c += 8) r.v[c / 8] = __ldcs(reinterpret_cast<const uint4*>(gact + elem_offset(TM, row, q * kCols + c)));
}
template <int kCols, int kPre>
__device__ __forceinline__ void epilogue_bwd(uint32_t tmem_acc, const uint8_t* __restrict__ gact, uint8_t* sdst, uint8_t* gdst, int g, int q, int lane,
                                             const ActRegs<kPre>& pre) {
  const int row = g * 32 + lane;
  const uint32_t taddr = tmem_acc + ((uint32_t)(g * 32) << 16) + (uint32_t)(q * kCols);
#pragma unroll
  for (int c = 0; c < kCols; c += 16) {
    uint32_t acc[16];
    tmem_ld16(taddr + c, acc);
    const int col = q * kCols + c;
    uint4 a_lo, a_hi;
    if (c < kPre) { a_lo = pre.v[c / 8]; a_hi = pre.v[c / 8 + 1]; }
    else {
      a_lo = __ldcs(reinterpret_cast<const uint4*>(gact + elem_offset(TM, row, col)));
      a_hi = __ldcs(reinterpret_cast<const uint4*>(gact + elem_offset(TM, row, col + 8)));
    }
    tmem_wait();
    const uint4 z_lo = dz_chunk(acc, a_lo), z_hi = dz_chunk(acc + 8, a_hi);
    if (sdst != nullptr) {
      *reinterpret_cast<uint4*>(sdst + elem_offset(TM, row, col)) = z_lo;
      *reinterpret_cast<uint4*>(sdst + elem_offset(TM, row, col + 8)) = z_hi;
    }
    __stcs(reinterpret_cast<uint4*>(gdst + elem_offset(TM, row, col)), z_lo);
    __stcs(reinterpret_cast<uint4*>(gdst + elem_offset(TM, row, col + 8)), z_hi);
  }
}

__global__ void __launch_bounds__(B_THREADS, 1) k_mlp_backward(const BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // count 1: bar_w (weight TMA), bar_acc3 / bar_acc2 / bar_acc1 (tcgen05.commit).  count 16: bar_dh, bar_z3, bar_z2 (operand written by
  // every epilogue warp).  Each completes once per tile: wait parity = CTA-local tile index & 1.
  __shared__ uint64_t bar_w, bar_dh, bar_acc3, bar_z3, bar_acc2, bar_z2, bar_acc1;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar_w, 1); mbar_init(&bar_dh, B_WARPS); mbar_init(&bar_acc3, 1); mbar_init(&bar_z3, B_WARPS); mbar_init(&bar_acc2, 1);
    mbar_init(&bar_z2, B_WARPS); mbar_init(&bar_acc1, 1);
    fence_mbar_init();
  }
  if (warp == B_WARPS) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int tile0 = blockIdx.x, stride = gridDim.x;
  const int n_loc = tile0 < a.num_tiles ? (a.num_tiles - tile0 + stride - 1) / stride : 0;
  uint8_t* DHs = smem + DH_OFF;
  uint8_t* DZ3s = smem + DZ3_OFF;
  uint8_t* DZ2s = smem + DZ2_OFF;

  if (warp == B_WARPS) {
    if (elect_one()) {
      mbar_expect_tx(&bar_w, WT_BYTES);
      tma_load_1d(smem, a.wt, WT_BYTES, &bar_w);
      mbar_wait(&bar_w, 0);
      const uint32_t sWh = smem_u32(smem + WHT_OFF), sW3 = smem_u32(smem + W3T_OFF), sW2 = smem_u32(smem + W2T_OFF);
      const uint32_t sDH = smem_u32(DHs), sZ3 = smem_u32(DZ3s), sZ2 = smem_u32(DZ2s);
      for (int it = 0; it < n_loc; ++it) {
        const uint32_t ph = it & 1;
        mbar_wait(&bar_dh, ph); tc_fence_after();
        issue_gemm<D3, DH>(tmem + BA3_COL, sDH, sWh); umma_commit(&bar_acc3);
        mbar_wait(&bar_z3, ph); tc_fence_after();
        issue_gemm<D2, D3>(tmem + BA2_COL, sZ3, sW3); umma_commit(&bar_acc2);
        mbar_wait(&bar_z2, ph); tc_fence_after();
        issue_gemm<D1, D2>(tmem + BA1_COL, sZ2, sW2); umma_commit(&bar_acc1);
      }
    }
  } else {
    const int g = warp & 3, q = warp >> 2;
    const int row = g * 32 + lane;
    for (int it = 0; it < n_loc; ++it) {
      const uint32_t ph = it & 1;
      const size_t tile = (size_t)(tile0 + it * stride);
      const int row0 = (int)tile * TM, rows = min(TM, a.n - row0);
      // ---- dH operand [128 x 16]: columns 0..11 dLoss/dmu, 12 dLoss/dvalue; rows past the batch are zero.  The previous tile's
      // GEMM on it completed before that tile's E3 (bar_acc3), so it may be overwritten.
      if (q == 0) {
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = 0.0f;
        if (row < rows) {
          const float4* gm = reinterpret_cast<const float4*>(a.grad_mu + (size_t)(row0 + row) * 12);
          const float4 m0 = __ldg(gm), m1 = __ldg(gm + 1), m2 = __ldg(gm + 2);
          v[0] = m0.x; v[1] = m0.y; v[2] = m0.z; v[3] = m0.w; v[4] = m1.x; v[5] = m1.y; v[6] = m1.z; v[7] = m1.w;
          v[8] = m2.x; v[9] = m2.y; v[10] = m2.z; v[11] = m2.w; v[12] = __ldg(a.grad_value + row0 + row);
        }
        const uint4 lo = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        const uint4 hi = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
        *reinterpret_cast<uint4*>(DHs + elem_offset(TM, row, 0)) = lo;
        *reinterpret_cast<uint4*>(DHs + elem_offset(TM, row, 8)) = hi;
        uint8_t* gdh = a.dh + tile * blob_bytes(DH);
        __stcs(reinterpret_cast<uint4*>(gdh + elem_offset(TM, row, 0)), lo);
        __stcs(reinterpret_cast<uint4*>(gdh + elem_offset(TM, row, 8)), hi);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_dh);
      // ---- E3: dZ3 = dA3 * ELU'(A3)
      ActRegs<16> p3;
      act_prefetch<D3 / 4, 16>(a.a3 + tile * blob_bytes(D3), g, q, lane, p3);
      mbar_wait(&bar_acc3, ph); tc_fence_after();
      epilogue_bwd<D3 / 4, 16>(tmem + BA3_COL, a.a3 + tile * blob_bytes(D3), DZ3s, a.dz3 + tile * blob_bytes(D3), g, q, lane, p3);
      fence_proxy_async(); tc_fence_before(); __syncwarp();
      if (lane == 0) mbar_arrive(&bar_z3);
      // ---- E2
      ActRegs<32> p2;
      act_prefetch<D2 / 4, 32>(a.a2 + tile * blob_bytes(D2), g, q, lane, p2);
      mbar_wait(&bar_acc2, ph); tc_fence_after();
      epilogue_bwd<D2 / 4, 32>(tmem + BA2_COL, a.a2 + tile * blob_bytes(D2), DZ2s, a.dz2 + tile * blob_bytes(D2), g, q, lane, p2);
      fence_proxy_async(); tc_fence_before(); __syncwarp();
      if (lane == 0) mbar_arrive(&bar_z2);
      // ---- E1: dZ1 leaves for the weight-gradient kernel only
      ActRegs<64> p1;
      act_prefetch<D1 / 4, 64>(a.a1 + tile * blob_bytes(D1), g, q, lane, p1);
      mbar_wait(&bar_acc1, ph); tc_fence_after();
      epilogue_bwd<D1 / 4, 64>(tmem + BA1_COL, a.a1 + tile * blob_bytes(D1), nullptr, a.dz1 + tile * blob_bytes(D1), g, q, lane, p1);
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == B_WARPS) tmem_dealloc(tmem, 512);
}

// nn.Linear W [O][I] -> K-major B operand of dA = dZ W: element (n = i, k = o).  All three transposed operands in one launch
// (the PPO update re-packs after every optimiser step).  Head: B[n = i (64)][k = o (16)]: o 0..11 = mean rows, 12 = value row,
// 13..15 zero.
constexpr int PT_WH = D3 * DH, PT_W3 = PT_WH + D3 * D2, PT_ALL = PT_W3 + D2 * D1;
__global__ void k_pack_t_all(const float* __restrict__ w2, const float* __restrict__ w3, const float* __restrict__ w_mu,
                             const float* __restrict__ w_value, uint8_t* __restrict__ blob_t) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < PT_WH) {
    const int o = idx / D3, i = idx - o * D3;
    const float v = o < 12 ? w_mu[idx] : (o == 12 ? w_value[i] : 0.0f);
    *reinterpret_cast<__nv_bfloat16*>(blob_t + WHT_OFF + elem_offset(D3, i, o)) = __float2bfloat16_rn(v);
  } else if (idx < PT_W3) {
    const int j = idx - PT_WH, o = j / D2, i = j - o * D2;
    *reinterpret_cast<__nv_bfloat16*>(blob_t + W3T_OFF + elem_offset(D2, i, o)) = __float2bfloat16_rn(w3[j]);
  } else if (idx < PT_ALL) {
    const int j = idx - PT_W3, o = j / D1, i = j - o * D1;
    *reinterpret_cast<__nv_bfloat16*>(blob_t + W2T_OFF + elem_offset(D1, i, o)) = __float2bfloat16_rn(w2[j]);
  }
}

// ---------------------------------------------------------------- weight gradients
// flat per-CTA partial layout (floats)
constexpr int P_W1 = 0;
constexpr int P_B1 = P_W1 + D1 * D0;
constexpr int P_W2 = P_B1 + D1;
constexpr int P_B2 = P_W2 + D2 * D1;
constexpr int P_W3 = P_B2 + D2;
constexpr int P_B3 = P_W3 + D3 * D2;
constexpr int P_WH = P_B3 + D3;          // [16][64]: rows 0..11 mean weights, 12 value weights
constexpr int P_BH = P_WH + DH * D3;
constexpr int P_COUNT = P_BH + DH;

// ring slot layouts (bytes)
constexpr int J1_DZ1 = 0, J1_A0 = blob_bytes(D1), J1_BYTES = J1_A0 + blob_bytes(D0);
constexpr int J2_DZ2 = 0, J2_A1 = blob_bytes(D2), J2_BYTES = J2_A1 + blob_bytes(D1);
constexpr int J3_DZ3 = 0, J3_A2 = blob_bytes(D3), J3_A3 = J3_A2 + blob_bytes(D2), J3_DH = J3_A3 + blob_bytes(D3), J3_BYTES = J3_DH + blob_bytes(DH);
constexpr int SLOT_BYTES = J2_BYTES;
static_assert(J1_BYTES <= SLOT_BYTES && J3_BYTES <= SLOT_BYTES, "slot");
// every operand read past its blob (M = 128 on a 64- or 16-wide blob) stays inside the slot: finite bf16 data of this or an earlier job
static_assert(J3_DH + 16 * lbo_bytes(TM) <= SLOT_BYTES && J3_A3 + 16 * lbo_bytes(TM) <= SLOT_BYTES && J3_DZ3 + 16 * lbo_bytes(TM) <= SLOT_BYTES, "overread");
constexpr int ONES_OFF = 2 * SLOT_BYTES;                    // [128 samples][8] bf16, column 0 = 1
constexpr int WG_SMEM_BYTES = ONES_OFF + TM * 16 + 64;
static_assert(WG_SMEM_BYTES <= 232448 - 1024, "exceeds 227 KB");
// TMEM columns
constexpr uint32_t C_W1A = 0, C_W1B = 64, C_W2 = 128, C_W3T = 384, C_B1A = 448, C_B1B = 456, C_B2 = 464, C_B3 = 472, C_WHT = 480, C_BH = 496;
constexpr int WG_THREADS = 160;                              // 4 read-out warps + the control warp

struct WgArgs {
  const uint8_t *dz1, *dz2, *dz3, *dh, *a0, *a1, *a2, *a3;   // blobs, tile-major
  float* partials;                                           // [grid][P_COUNT]
  int num_tiles;
};

// MN-major view of a blob: SBO (MN groups of 8 features) = 2064 B, LBO (K groups of 8 samples) = 128 B
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((128 >> 4) & 0x3FFF) << 16) | ((uint64_t)((lbo_bytes(TM) >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t instr_desc_mn(int m, int n) {      // bf16 x bf16 -> f32, A and B MN-major (bits 15, 16)
  return instr_desc(m, n) | (1u << 15) | (1u << 16);
}
// D[M = 128 features of a, N features of b] (+)= sum over the tile's 128 samples: 8 K steps of 16 samples (256 B apart in both blobs)
template <int kN>
__device__ __forceinline__ void wgrad_mma(uint32_t acc, uint32_t a_saddr, uint32_t b_saddr, uint32_t accumulate) {
  constexpr uint32_t idesc = instr_desc_mn(TM, kN);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    umma_bf16(acc, smem_desc_mn(a_saddr + ks * 256), smem_desc_mn(b_saddr + ks * 256), idesc, (accumulate | (uint32_t)(ks > 0)));
}

__global__ void __launch_bounds__(WG_THREADS, 1) k_mlp_wgrad(const WgArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full[2], bar_free[2], bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar_full[0], 1); mbar_init(&bar_full[1], 1); mbar_init(&bar_free[0], 1); mbar_init(&bar_free[1], 1); mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  // the ones operand: [s][8] bf16 with column 0 = 1 (2048 contiguous bytes in either view)
  for (int i = tid; i < TM * 4; i += WG_THREADS) reinterpret_cast<uint32_t*>(smem + ONES_OFF)[i] = (i & 3) == 0 ? 0x00003F80u : 0u;
  fence_proxy_async();
  if (warp == 4) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int tile0 = blockIdx.x, stride = gridDim.x;
  const int n_loc = tile0 < a.num_tiles ? (a.num_tiles - tile0 + stride - 1) / stride : 0;
  const int n_jobs = 3 * n_loc;

  if (warp == 4) {
    if (elect_one()) {
      const uint32_t sOnes = smem_u32(smem + ONES_OFF);
      auto load = [&](int j) {                       // job j = (tile, which) -> ring slot j & 1
        const int it = j / 3, which = j - 3 * it, slot = j & 1;
        const size_t t = (size_t)(tile0 + it * stride);
        uint8_t* s = smem + slot * SLOT_BYTES;
        uint64_t* bar = &bar_full[slot];
        if (which == 0) {
          mbar_expect_tx(bar, J1_BYTES);
          tma_load_1d(s + J1_DZ1, a.dz1 + t * blob_bytes(D1), blob_bytes(D1), bar);
          tma_load_1d(s + J1_A0, a.a0 + t * blob_bytes(D0), blob_bytes(D0), bar);
        } else if (which == 1) {
          mbar_expect_tx(bar, J2_BYTES);
          tma_load_1d(s + J2_DZ2, a.dz2 + t * blob_bytes(D2), blob_bytes(D2), bar);
          tma_load_1d(s + J2_A1, a.a1 + t * blob_bytes(D1), blob_bytes(D1), bar);
        } else {
          mbar_expect_tx(bar, J3_BYTES);
          tma_load_1d(s + J3_DZ3, a.dz3 + t * blob_bytes(D3), blob_bytes(D3), bar);
          tma_load_1d(s + J3_A2, a.a2 + t * blob_bytes(D2), blob_bytes(D2), bar);
          tma_load_1d(s + J3_A3, a.a3 + t * blob_bytes(D3), blob_bytes(D3), bar);
          tma_load_1d(s + J3_DH, a.dh + t * blob_bytes(DH), blob_bytes(DH), bar);
        }
      };
      if (n_jobs > 0) load(0);
      for (int j = 0; j < n_jobs; ++j) {
        const int slot = j & 1, use = j >> 1;               // use-th use of this slot
        if (j + 1 < n_jobs) {
          const int nslot = (j + 1) & 1, nuse = (j + 1) >> 1;
          if (nuse > 0) mbar_wait(&bar_free[nslot], (uint32_t)((nuse - 1) & 1));   // the MMAs of the slot's previous job have drained
          load(j + 1);
        }
        mbar_wait(&bar_full[slot], (uint32_t)(use & 1));
        tc_fence_after();
        const uint32_t s = smem_u32(smem + slot * SLOT_BYTES);
        const uint32_t acc = j >= 3 ? 1u : 0u;              // first tile of the CTA initialises the accumulators
        const int which = j % 3;
        if (which == 0) {
          wgrad_mma<D0>(tmem + C_W1A, s + J1_DZ1, s + J1_A0, acc);
          wgrad_mma<D0>(tmem + C_W1B, s + J1_DZ1 + 16 * lbo_bytes(TM), s + J1_A0, acc);
          wgrad_mma<8>(tmem + C_B1A, s + J1_DZ1, sOnes, acc);
          wgrad_mma<8>(tmem + C_B1B, s + J1_DZ1 + 16 * lbo_bytes(TM), sOnes, acc);
        } else if (which == 1) {
          wgrad_mma<D1>(tmem + C_W2, s + J2_DZ2, s + J2_A1, acc);
          wgrad_mma<8>(tmem + C_B2, s + J2_DZ2, sOnes, acc);
        } else {
          wgrad_mma<D3>(tmem + C_W3T, s + J3_A2, s + J3_DZ3, acc);
          wgrad_mma<8>(tmem + C_B3, s + J3_DZ3, sOnes, acc);
          wgrad_mma<DH>(tmem + C_WHT, s + J3_A3, s + J3_DH, acc);
          wgrad_mma<8>(tmem + C_BH, s + J3_DH, sOnes, acc);
        }
        umma_commit(&bar_free[slot]);
      }
      umma_commit(&bar_done);
    }
  } else if (n_loc > 0) {
    // ================================================================== read-out: thread = accumulator row
    mbar_wait(&bar_done, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const uint32_t t0 = tmem + ((uint32_t)(warp * 32) << 16);
    float* P = a.partials + (size_t)blockIdx.x * P_COUNT;
    uint32_t r[32];
    auto rows_out = [&](uint32_t col0, int ncols, float* dst /* row-major, `ncols` per row */) {
      for (int c = 0; c < ncols; c += 32) {
        tmem_ld32(t0 + col0 + c, r);
        tmem_wait();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          reinterpret_cast<float4*>(dst + (size_t)row * ncols + c)[k] =
              make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), __uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]));
      }
    };
    rows_out(C_W1A, D0, P + P_W1);                                   // out rows 0..127
    rows_out(C_W1B, D0, P + P_W1 + 128 * D0);                        // out rows 128..255
    rows_out(C_W2, D1, P + P_W2);
    {                                                                // dW3^T: row = in, column = out -> W3[out][in]
      for (int c = 0; c < D3; c += 32) {
        tmem_ld32(t0 + C_W3T + c, r);
        tmem_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) P[P_W3 + (c + k) * D2 + row] = __uint_as_float(r[k]);
      }
    }
    uint32_t b[8];
    tmem_ld8(t0 + C_B1A, b); tmem_wait(); P[P_B1 + row] = __uint_as_float(b[0]);
    tmem_ld8(t0 + C_B1B, b); tmem_wait(); P[P_B1 + 128 + row] = __uint_as_float(b[0]);
    tmem_ld8(t0 + C_B2, b); tmem_wait(); P[P_B2 + row] = __uint_as_float(b[0]);
    tmem_ld8(t0 + C_B3, b); tmem_wait(); if (row < D3) P[P_B3 + row] = __uint_as_float(b[0]);
    {                                                                // dWh^T: row = in (< 64), column = o -> Wh[o][in]
      uint32_t h[16];
      tmem_ld16(t0 + C_WHT, h);
      tmem_wait();
      if (row < D3) {
#pragma unroll
        for (int o = 0; o < DH; ++o) P[P_WH + o * D3 + row] = __uint_as_float(h[o]);
      }
    }
    tmem_ld8(t0 + C_BH, b); tmem_wait(); if (row < DH) P[P_BH + row] = __uint_as_float(b[0]);
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

// sum of the per-CTA partials in CTA order -> the parameters' .grad tensors (nn.Linear layouts)
struct ReduceArgs {
  const float* partials; int num_ctas;
  float *w1, *b1, *w2, *b2, *w3, *b3, *w_mu, *b_mu, *w_value, *b_value;
};
// flat partial index j -> where that gradient element lives (NULL: padding rows of the head operand)
__device__ __forceinline__ float* wgrad_dst(const ReduceArgs& a, int j) {
  if (j < P_B1) return a.w1 + (j - P_W1);
  if (j < P_W2) return a.b1 + (j - P_B1);
  if (j < P_B2) return a.w2 + (j - P_W2);
  if (j < P_W3) return a.b2 + (j - P_B2);
  if (j < P_B3) return a.w3 + (j - P_W3);
  if (j < P_WH) return a.b3 + (j - P_B3);
  if (j < P_BH) {
    const int k = j - P_WH, o = k / D3;
    return o < 12 ? a.w_mu + k : (o == 12 ? a.w_value + (k - 12 * D3) : nullptr);
  }
  const int o = j - P_BH;
  return o < 12 ? a.b_mu + o : (o == 12 ? a.b_value : nullptr);
}
__device__ __forceinline__ float partial_sum(const ReduceArgs& a, int j) {
  float s = 0.0f;
  int c = 0;
  for (; c + 8 <= a.num_ctas; c += 8) {                       // eight independent loads in flight; summed in CTA order
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(a.partials + (size_t)(c + u) * P_COUNT + j);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; c < a.num_ctas; ++c) s += __ldcg(a.partials + (size_t)c * P_COUNT + j);
  return s;
}
__global__ void k_wgrad_reduce(const ReduceArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P_COUNT) return;
  float* dst = wgrad_dst(a, j);
  if (dst != nullptr) *dst = partial_sum(a, j);
}

// The same reduction fused with the data-parallel gradient all-reduce (SURVEY 8f-1 "fused with the grad allreduce"): every
// gradient element is pushed into the peers' receive slots the moment its partial sum exists (NVLink stores, lm_comm.cuh), the
// ranks exchange one flag, and each rank sums the slots in rank order straight into its gradient arena.  `arena` is the flat
// fp32 buffer all gradient tensors are views of; the elements wgrad does not produce (log_std gradient, the loss head's
// statistics in the tail: two ranges) travel in the same message.
struct ReduceCommArgs {
  ReduceArgs r;
  comm::Ctx c;
  float* arena; long long arena_n;
  long long extra_off[2], extra_n[2];
};
__global__ void __launch_bounds__(1024, 1) k_wgrad_reduce_allreduce(const ReduceCommArgs a) {
  const unsigned int epoch = comm::epoch_of(a.c);
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;
  for (int j = gtid; j < P_COUNT; j += gstride) {
    float* dst = wgrad_dst(a.r, j);
    if (dst != nullptr) comm::push(a.c, epoch, (long long)(dst - a.arena), partial_sum(a.r, j));
  }
#pragma unroll
  for (int x = 0; x < 2; ++x)
    for (long long i = gtid; i < a.extra_n[x]; i += gstride) comm::push(a.c, epoch, a.extra_off[x] + i, a.arena[a.extra_off[x] + i]);
  if (comm::signal_and_wait(a.c, epoch)) {
    for (long long f = gtid; f < a.arena_n; f += gstride) a.arena[f] = comm::sum(a.c, epoch, f);
  }
  comm::leave(a.c);
}

}  // namespace mlp_train
}  // namespace lm

using namespace lm;
using namespace lm::mlp_train;


int lm_mlp_train_pack(LmMlp* m, const float* w2, const float* w3, const float* w_mu, const float* w_value, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!m->blob_t) {
    if (cudaMalloc(&m->blob_t, WT_BYTES) != cudaSuccess) { m->blob_t = nullptr; return lm_fail((int)cudaErrorMemoryAllocation, "lm_mlp_pack: cudaMalloc failed"); }
    cudaMemsetAsync(m->blob_t, 0, WT_BYTES, s);           // padding bytes once; every element is rewritten by each pack
  }
  k_pack_t_all<<<(PT_ALL + 255) / 256, 256, 0, s>>>(w2, w3, w_mu, w_value, m->blob_t);
  return lm_check_launch("lm_mlp_pack (transposed operands)");
}

void lm_mlp_train_release(LmMlp* m) {
  if (m->blob_t) cudaFree(m->blob_t);
  m->blob_t = nullptr;
}

extern "C" {

int64_t lm_mlp_blob_bytes(int64_t rows, int32_t d) { return ((rows + TM - 1) / TM) * (int64_t)blob_bytes(d); }

int lm_mlp_blob_pack(const float* x, void* blob, int64_t rows, int32_t d, void* stream) {
  if (!x || !blob || rows <= 0 || d <= 0 || (d & 7)) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_blob_pack: bad argument (d must be a multiple of 8)");
  const long long total = ((rows + TM - 1) / TM) * TM * d;
  k_blob_pack<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, (uint8_t*)blob, rows, d);
  return lm_check_launch("lm_mlp_blob_pack");
}

int lm_mlp_blob_unpack(const void* blob, float* x, int64_t rows, int32_t d, void* stream) {
  if (!x || !blob || rows <= 0 || d <= 0 || (d & 7)) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_blob_unpack: bad argument (d must be a multiple of 8)");
  k_blob_unpack<<<(unsigned)((rows * d + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)blob, x, rows, d);
  return lm_check_launch("lm_mlp_blob_unpack");
}

int lm_mlp_wgrad(const void* dz1, const void* dz2, const void* dz3, const void* dh, const void* a0, const void* a1, const void* a2,
                 const void* a3, int64_t rows, float* workspace, int32_t workspace_ctas, float* g_w1, float* g_b1, float* g_w2, float* g_b2,
                 float* g_w3, float* g_b3, float* g_w_mu, float* g_b_mu, float* g_w_value, float* g_b_value, void* stream) {
  if (!dz1 || !dz2 || !dz3 || !dh || !a0 || !a1 || !a2 || !a3 || !workspace || rows <= 0 || workspace_ctas <= 0)
    return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_wgrad: bad argument");
  if (!g_w1 || !g_b1 || !g_w2 || !g_b2 || !g_w3 || !g_b3 || !g_w_mu || !g_b_mu || !g_w_value || !g_b_value)
    return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_wgrad: NULL gradient tensor");
  {   // per call and per current device: the entry takes no handle, and the attribute is a cheap host-side setting
    cudaError_t e = cudaFuncSetAttribute(k_mlp_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
    if (e != cudaSuccess) return lm_fail((int)e, "lm_mlp_wgrad: smem attribute: %s", cudaGetErrorString(e));
  }
  WgArgs a;
  a.dz1 = (const uint8_t*)dz1; a.dz2 = (const uint8_t*)dz2; a.dz3 = (const uint8_t*)dz3; a.dh = (const uint8_t*)dh;
  a.a0 = (const uint8_t*)a0; a.a1 = (const uint8_t*)a1; a.a2 = (const uint8_t*)a2; a.a3 = (const uint8_t*)a3;
  a.partials = workspace; a.num_tiles = (int)((rows + TM - 1) / TM);
  const int grid = a.num_tiles < workspace_ctas ? a.num_tiles : workspace_ctas;
  k_mlp_wgrad<<<grid, WG_THREADS, WG_SMEM_BYTES, (cudaStream_t)stream>>>(a);
  ReduceArgs r{workspace, grid, g_w1, g_b1, g_w2, g_b2, g_w3, g_b3, g_w_mu, g_b_mu, g_w_value, g_b_value};
  k_wgrad_reduce<<<(P_COUNT + 255) / 256, 256, 0, (cudaStream_t)stream>>>(r);
  return lm_check_launch("lm_mlp_wgrad");
}

int lm_mlp_wgrad_allreduce(LmComm* comm_h, const void* dz1, const void* dz2, const void* dz3, const void* dh, const void* a0, const void* a1,
                           const void* a2, const void* a3, int64_t rows, float* workspace, int32_t workspace_ctas, float* g_w1, float* g_b1,
                           float* g_w2, float* g_b2, float* g_w3, float* g_b3, float* g_w_mu, float* g_b_mu, float* g_w_value, float* g_b_value,
                           float* arena, int64_t arena_floats, int64_t extra_off0, int64_t extra_n0, int64_t extra_off1, int64_t extra_n1,
                           void* stream) {
  if (!comm_h || !arena || arena_floats <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_wgrad_allreduce: bad argument");
  if (!dz1 || !dz2 || !dz3 || !dh || !a0 || !a1 || !a2 || !a3 || !workspace || rows <= 0 || workspace_ctas <= 0)
    return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_wgrad_allreduce: bad argument");
  float* g[10] = {g_w1, g_b1, g_w2, g_b2, g_w3, g_b3, g_w_mu, g_b_mu, g_w_value, g_b_value};
  const long long gn[10] = {D1 * D0, D1, D2 * D1, D2, D3 * D2, D3, 12 * D3, 12, D3, 1};
  long long covered = (extra_n0 > 0 ? extra_n0 : 0) + (extra_n1 > 0 ? extra_n1 : 0);
  for (int i = 0; i < 10; ++i) {
    if (!g[i] || g[i] < arena || g[i] + gn[i] > arena + arena_floats)
      return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_wgrad_allreduce: gradient tensor %d is not a view of the arena", i);
    covered += gn[i];
  }
  if (covered != arena_floats) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_wgrad_allreduce: arena of %lld floats, %lld accounted for", (long long)arena_floats, covered);
  if (arena_floats > lm_comm_slot_floats(comm_h)) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_wgrad_allreduce: arena exceeds the communicator's slot size");
  {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
    if (e != cudaSuccess) return lm_fail((int)e, "lm_mlp_wgrad_allreduce: smem attribute: %s", cudaGetErrorString(e));
  }
  WgArgs a;
  a.dz1 = (const uint8_t*)dz1; a.dz2 = (const uint8_t*)dz2; a.dz3 = (const uint8_t*)dz3; a.dh = (const uint8_t*)dh;
  a.a0 = (const uint8_t*)a0; a.a1 = (const uint8_t*)a1; a.a2 = (const uint8_t*)a2; a.a3 = (const uint8_t*)a3;
  a.partials = workspace; a.num_tiles = (int)((rows + TM - 1) / TM);
  const int grid = a.num_tiles < workspace_ctas ? a.num_tiles : workspace_ctas;
  k_mlp_wgrad<<<grid, WG_THREADS, WG_SMEM_BYTES, (cudaStream_t)stream>>>(a);
  ReduceCommArgs rc;
  rc.r = ReduceArgs{workspace, grid, g_w1, g_b1, g_w2, g_b2, g_w3, g_b3, g_w_mu, g_b_mu, g_w_value, g_b_value};
  rc.c = *lm_comm_ctx(comm_h);
  rc.arena = arena; rc.arena_n = arena_floats;
  rc.extra_off[0] = extra_off0; rc.extra_n[0] = extra_n0 > 0 ? extra_n0 : 0; rc.extra_off[1] = extra_off1; rc.extra_n[1] = extra_n1 > 0 ? extra_n1 : 0;
  const int want = (P_COUNT + 1023) / 1024, cap = lm_comm_max_ctas(comm_h);
  k_wgrad_reduce_allreduce<<<want < cap ? want : cap, 1024, 0, (cudaStream_t)stream>>>(rc);
  return lm_check_launch("lm_mlp_wgrad_allreduce");
}

int32_t lm_mlp_wgrad_workspace_floats(void) { return P_COUNT; }

int lm_mlp_backward(LmMlp* m, const float* grad_mu, const float* grad_value, const void* a1, const void* a2, const void* a3, void* dh, void* dz3,
                    void* dz2, void* dz1, int32_t n, void* stream) {
  if (!m || !grad_mu || !grad_value || !a1 || !a2 || !a3 || !dh || !dz3 || !dz2 || !dz1 || n <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_backward: bad argument");
  if (!m->packed || !m->blob_t) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_backward: weights not packed (call lm_mlp_pack)");
  if ((((uintptr_t)grad_mu) | ((uintptr_t)a1) | ((uintptr_t)a2) | ((uintptr_t)a3) | ((uintptr_t)dh) | ((uintptr_t)dz3) | ((uintptr_t)dz2) | ((uintptr_t)dz1)) & 15)
    return lm_fail(LM_ERR_ALIGNMENT, "lm_mlp_backward: grad_mu / blobs must be 16-byte aligned");
  {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM_BYTES);
    if (e != cudaSuccess) return lm_fail((int)e, "lm_mlp_backward: smem attribute: %s", cudaGetErrorString(e));
  }
  BwdArgs a;
  a.wt = m->blob_t; a.grad_mu = grad_mu; a.grad_value = grad_value;
  a.a1 = (const uint8_t*)a1; a.a2 = (const uint8_t*)a2; a.a3 = (const uint8_t*)a3;
  a.dh = (uint8_t*)dh; a.dz3 = (uint8_t*)dz3; a.dz2 = (uint8_t*)dz2; a.dz1 = (uint8_t*)dz1;
  a.n = n; a.num_tiles = (n + TM - 1) / TM;
  const int grid = a.num_tiles < m->sm_count ? a.num_tiles : m->sm_count;
  k_mlp_backward<<<grid, B_THREADS, BWD_SMEM_BYTES, (cudaStream_t)stream>>>(a);
  return lm_check_launch("lm_mlp_backward");
}


}  // extern "C"
