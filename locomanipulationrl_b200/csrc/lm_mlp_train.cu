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

#include "lm_common.cuh"
#include "lm_host.h"
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
__global__ void k_wgrad_reduce(const ReduceArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P_COUNT) return;
  float s = 0.0f;
  for (int c = 0; c < a.num_ctas; ++c) s += a.partials[(size_t)c * P_COUNT + j];
  if (j < P_B1) a.w1[j - P_W1] = s;
  else if (j < P_W2) a.b1[j - P_B1] = s;
  else if (j < P_B2) a.w2[j - P_W2] = s;
  else if (j < P_W3) a.b2[j - P_B2] = s;
  else if (j < P_B3) a.w3[j - P_W3] = s;
  else if (j < P_WH) a.b3[j - P_B3] = s;
  else if (j < P_BH) {
    const int k = j - P_WH, o = k / D3;
    if (o < 12) a.w_mu[k] = s; else if (o == 12) a.w_value[k - 12 * D3] = s;
  } else {
    const int o = j - P_BH;
    if (o < 12) a.b_mu[o] = s; else if (o == 12) a.b_value[0] = s;
  }
}

}  // namespace mlp_train
}  // namespace lm

using namespace lm;
using namespace lm::mlp_train;

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
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
    if (e != cudaSuccess) return lm_fail((int)e, "lm_mlp_wgrad: smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
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

int32_t lm_mlp_wgrad_workspace_floats(void) { return P_COUNT; }

}  // extern "C"
