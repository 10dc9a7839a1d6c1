// MLP actor-critic forward on 5th-gen tensor cores (tcgen05 + TMEM), persistent, weights resident.
//
// Network (scripts/skrl_ppo_locomanipulation.py:33-42,57-61; rl_games actor_critic from
// cfg/train/JointLocomanipulationPPO.yaml:10-32):
//   x[64] -> Linear(64,256) ELU -> Linear(256,128) ELU -> Linear(128,64) ELU -> { mean Linear(64,12), value Linear(64,1) }
//
// One CTA per SM, 512 threads (16 epilogue warps: 4 TMEM lane groups x 4 column quarters).  All four weight matrices (bf16, 115 KB) are pulled into shared memory
// once per CTA by a TMA bulk copy and stay there; the CTA then walks 128-env tiles:
//   TMA bulk load of the fp32 obs tile -> normalise/clamp -> bf16 A operand in the UMMA canonical
//   K-major layout -> tcgen05.mma (M=128, N=256/128/64/16, fp32 accumulators in TMEM) ->
//   tcgen05.ld epilogue (+bias, ELU, bf16) writes the next layer's A operand straight back to shared
//   memory.  Activations never touch HBM; per env the kernel reads 256 B and writes 52 B.
// The next tile's obs load is issued as soon as its staging region is free, so it overlaps the
// last two layers of the current tile.
#include <cuda_bf16.h>

#include "lm_common.cuh"
#include "lm_host.h"

namespace lm {
namespace mlp {

constexpr int TM = 128;                 // rows (envs) per tile == UMMA M
constexpr int THREADS = 512;              // 16 warps: 4 TMEM lane groups x 4 column quarters
constexpr int D0 = 64, D1 = 256, D2 = 128, D3 = 64, DH = 16;   // head: 12 mean + 1 value + 3 pad

// UMMA canonical K-major, no swizzle: core matrix = 8 rows x 16 B, rows 16 B apart; core matrices
// 128 B apart along M/N (SBO); K chunks (8 bf16) `lbo` bytes apart.  lbo = rows*16 + 16: the extra
// 16 B rotates banks so writing 8 chunks of one row is conflict free.
__host__ __device__ constexpr int lbo_bytes(int rows) { return rows * 16 + 16; }
__host__ __device__ constexpr int operand_bytes(int rows, int k) { return (k / 8) * lbo_bytes(rows); }
__host__ __device__ constexpr int elem_offset(int rows, int row, int k) {   // byte offset of element (row, k)
  return (k / 8) * lbo_bytes(rows) + row * 16 + (k % 8) * 2;
}

// packed parameter blob (global, produced by lm_mlp_pack; mirrored 1:1 in shared memory)
constexpr int W1_OFF = 0;
constexpr int W2_OFF = W1_OFF + operand_bytes(D1, D0);
constexpr int W3_OFF = W2_OFF + operand_bytes(D2, D1);
constexpr int WH_OFF = W3_OFF + operand_bytes(D3, D2);
constexpr int BIAS_OFF = WH_OFF + operand_bytes(DH, D3);          // fp32 biases: 256 + 128 + 64 + 16
constexpr int BLOB_BYTES = BIAS_OFF + (D1 + D2 + D3 + DH) * 4;
static_assert(BLOB_BYTES % 16 == 0 && W2_OFF % 16 == 0 && W3_OFF % 16 == 0 && WH_OFF % 16 == 0 && BIAS_OFF % 16 == 0, "align");

// activation regions
constexpr int RA_BYTES = operand_bytes(TM, D2);                   // A2 [128 x 128]; also the fp32 obs stage (32 KB)
constexpr int RB_BYTES = operand_bytes(TM, D1);                   // A1 [128 x 256]; also A0, A3
static_assert(RA_BYTES >= TM * D0 * 4, "obs stage must fit region A");
constexpr int RA_OFF = BLOB_BYTES;
constexpr int RB_OFF = RA_OFF + RA_BYTES;
constexpr int NORM_OFF = RB_OFF + RB_BYTES;                       // mean[64], inv_std[64] fp32
constexpr int SMEM_BYTES = NORM_OFF + 2 * D0 * 4 + 64;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");

// ---------------------------------------------------------------- tcgen05 wrappers
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, int lbo) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout NONE
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((128 >> 4) & 0x3FFF) << 32) |
         (1ull << 46);
}
__host__ __device__ constexpr uint32_t instr_desc(int m, int n) {
  // InstrDescriptor: c_format F32 (1) [4,6), a/b format BF16 (1) [7,10)/[10,13), K-major both, n>>3 [17,23), m>>4 [24,29)
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// 32 lanes x 16 / 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i).  The loads are
// asynchronous; tmem_wait() must precede the first use of the registers.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t r[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float elu(float x) { return x > 0.0f ? x : __expf(x) - 1.0f; }   // nn.ELU(alpha=1)
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// issue the K/16 MMAs of one layer (single thread) and commit to `bar`
template <int kN, int kK>
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_saddr, uint32_t b_saddr, uint64_t* bar) {
  tc_fence_after();
  constexpr uint32_t idesc = instr_desc(TM, kN);
#pragma unroll
  for (int ks = 0; ks < kK / 16; ++ks) {
    const uint64_t ad = smem_desc(a_saddr + ks * 2 * lbo_bytes(TM), lbo_bytes(TM));
    const uint64_t bd = smem_desc(b_saddr + ks * 2 * lbo_bytes(kN), lbo_bytes(kN));
    umma_bf16(tmem_d, ad, bd, idesc, ks > 0 ? 1u : 0u);
  }
  umma_commit(bar);
}

// 8 accumulator columns -> +bias, ELU, bf16 -> one 16-byte K chunk of the next A operand
__device__ __forceinline__ void emit_chunk(const uint32_t* acc, const float* __restrict__ bias, uint8_t* dst) {
  const float4 b0 = *reinterpret_cast<const float4*>(bias), b1 = *reinterpret_cast<const float4*>(bias + 4);
  const float e0 = elu(__uint_as_float(acc[0]) + b0.x), e1 = elu(__uint_as_float(acc[1]) + b0.y);
  const float e2 = elu(__uint_as_float(acc[2]) + b0.z), e3 = elu(__uint_as_float(acc[3]) + b0.w);
  const float e4 = elu(__uint_as_float(acc[4]) + b1.x), e5 = elu(__uint_as_float(acc[5]) + b1.y);
  const float e6 = elu(__uint_as_float(acc[6]) + b1.z), e7 = elu(__uint_as_float(acc[7]) + b1.w);
  *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(e0, e1), pack_bf16(e2, e3), pack_bf16(e4, e5), pack_bf16(e6, e7));
}

// hidden-layer epilogue: TMEM fp32 [128 x kN] -> +bias, ELU, bf16 -> next A operand (K = kN) in `dst`.
// Warp w owns TMEM lanes 32*(w&3).. and the column quarter (w>>2); all of its columns are fetched with one or two
// asynchronous tcgen05.ld before the single wait, so the TMEM latency is paid once per layer.
template <int kN>
__device__ __forceinline__ void epilogue_hidden(uint32_t tmem_base, const float* __restrict__ bias, uint8_t* dst, int warp, int lane) {
  constexpr int PER = kN / 4;                       // columns per warp: 64, 32 or 16
  const int row = (warp & 3) * 32 + lane;
  const int c0 = (warp >> 2) * PER;
  const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
  if constexpr (PER == 64) {
    uint32_t a[32], b[32];
    tmem_ld32(taddr, a);
    tmem_ld32(taddr + 32, b);
    tmem_wait();
#pragma unroll
    for (int k = 0; k < 4; ++k) emit_chunk(a + 8 * k, bias + c0 + 8 * k, dst + elem_offset(TM, row, c0 + 8 * k));
#pragma unroll
    for (int k = 0; k < 4; ++k) emit_chunk(b + 8 * k, bias + c0 + 32 + 8 * k, dst + elem_offset(TM, row, c0 + 32 + 8 * k));
  } else if constexpr (PER == 32) {
    uint32_t a[32];
    tmem_ld32(taddr, a);
    tmem_wait();
#pragma unroll
    for (int k = 0; k < 4; ++k) emit_chunk(a + 8 * k, bias + c0 + 8 * k, dst + elem_offset(TM, row, c0 + 8 * k));
  } else {
    uint32_t a[16];
    tmem_ld16(taddr, a);
    tmem_wait();
#pragma unroll
    for (int k = 0; k < 2; ++k) emit_chunk(a + 8 * k, bias + c0 + 8 * k, dst + elem_offset(TM, row, c0 + 8 * k));
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
};

__global__ void __launch_bounds__(THREADS, 1) k_mlp_forward(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_w, bar_obs, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* RA = smem + RA_OFF;
  uint8_t* RB = smem + RB_OFF;
  float* norm = reinterpret_cast<float*>(smem + NORM_OFF);
  const float* bias1 = reinterpret_cast<const float*>(smem + BIAS_OFF);
  const float* bias2 = bias1 + D1;
  const float* bias3 = bias2 + D2;
  const float* biash = bias3 + D3;

  if (tid == 0) {
    mbar_init(&bar_w, 1); mbar_init(&bar_obs, 1); mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  if (tid < D0) {
    norm[tid] = a.norm_mean ? a.norm_mean[tid] : 0.0f;
    norm[D0 + tid] = a.norm_inv_std ? a.norm_inv_std[tid] : 1.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  int tile = blockIdx.x;
  if (tid == 0) {
    // resident weights + biases: one bulk copy per CTA
    mbar_expect_tx(&bar_w, BLOB_BYTES);
    tma_load_1d(smem, a.blob, BLOB_BYTES, &bar_w);
    if (tile < a.num_tiles) {
      const int rows = min(TM, a.n - tile * TM);
      mbar_expect_tx(&bar_obs, (uint32_t)rows * D0 * 4);
      tma_load_1d(RA, a.obs + (size_t)tile * TM * D0, (uint32_t)rows * D0 * 4, &bar_obs);
    }
  }
  mbar_wait(&bar_w, 0);
  uint32_t ph_obs = 0, ph_mma = 0;
  const bool has_norm = a.norm_mean != nullptr;

  for (; tile < a.num_tiles; tile += gridDim.x) {
    const int row0 = tile * TM;
    const int rows = min(TM, a.n - row0);
    // ---- obs tile fp32 (region A) -> normalise/clamp -> bf16 A0 (region B)
    mbar_wait(&bar_obs, ph_obs); ph_obs ^= 1;
    for (int i = tid; i < TM * (D0 / 8); i += THREADS) {
      const int row = i >> 3, c = i & 7;
      float x[8];
      if (row < rows) {
        const float4 p = *reinterpret_cast<const float4*>(RA + row * (D0 * 4) + c * 32);
        const float4 q = *reinterpret_cast<const float4*>(RA + row * (D0 * 4) + c * 32 + 16);
        x[0] = p.x; x[1] = p.y; x[2] = p.z; x[3] = p.w; x[4] = q.x; x[5] = q.y; x[6] = q.z; x[7] = q.w;
        if (has_norm) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float y = (x[j] - norm[c * 8 + j]) * norm[D0 + c * 8 + j];
            x[j] = fminf(fmaxf(y, -a.norm_clip), a.norm_clip);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = 0.0f;
      }
      *reinterpret_cast<uint4*>(RB + elem_offset(TM, row, c * 8)) =
          make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- layer 1: [128x64] x W1^T -> TMEM[128x256]
    if (tid == 0) issue_layer<D1, D0>(tmem, smem_u32(RB), smem_u32(smem + W1_OFF), &bar_mma);
    mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1;
    tc_fence_after();
    epilogue_hidden<D1>(tmem, bias1, RB, warp, lane);          // A1 overwrites A0 (layer-1 MMA has completed)
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- layer 2: [128x256] x W2^T -> TMEM[128x128]
    if (tid == 0) issue_layer<D2, D1>(tmem, smem_u32(RB), smem_u32(smem + W2_OFF), &bar_mma);
    mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1;
    tc_fence_after();
    epilogue_hidden<D2>(tmem, bias2, RA, warp, lane);          // A2 -> region A (obs stage already consumed)
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- layer 3: [128x128] x W3^T -> TMEM[128x64]
    if (tid == 0) issue_layer<D3, D2>(tmem, smem_u32(RA), smem_u32(smem + W3_OFF), &bar_mma);
    mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1;
    tc_fence_after();
    // region A is free again: prefetch the next tile's obs while layers 3-4 finish
    const int next = tile + gridDim.x;
    if (tid == 0 && next < a.num_tiles) {
      const int nrows = min(TM, a.n - next * TM);
      mbar_expect_tx(&bar_obs, (uint32_t)nrows * D0 * 4);
      tma_load_1d(RA, a.obs + (size_t)next * TM * D0, (uint32_t)nrows * D0 * 4, &bar_obs);
    }
    epilogue_hidden<D3>(tmem, bias3, RB, warp, lane);          // A3 -> region B
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- heads: [128x64] x Wh^T -> TMEM[128x16]  (cols 0..11 mean, 12 value)
    if (tid == 0) issue_layer<DH, D3>(tmem, smem_u32(RB), smem_u32(smem + WH_OFF), &bar_mma);
    mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1;
    tc_fence_after();
    if (warp < 4) {
      uint32_t acc[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), acc);
      tmem_wait();
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
      const int row = warp * 32 + lane;
      if (row < rows) {
        float* m = a.mu + (size_t)(row0 + row) * 12;
        float o[13];
#pragma unroll
        for (int j = 0; j < 13; ++j) o[j] = v[j] + biash[j];
        reinterpret_cast<float4*>(m)[0] = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4*>(m)[1] = make_float4(o[4], o[5], o[6], o[7]);
        reinterpret_cast<float4*>(m)[2] = make_float4(o[8], o[9], o[10], o[11]);
        a.value[row0 + row] = o[12];
      }
    }
    tc_fence_before();
    __syncthreads();   // all TMEM reads done before the next tile's layer-1 MMA overwrites the accumulator
  }
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// fp32 nn.Linear [out,in] -> bf16 canonical operand; rows >= out are zero
__global__ void k_pack_operand(const float* __restrict__ w, uint8_t* __restrict__ dst, int rows_total, int rows_valid, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows_total * k) return;
  const int row = i / k, col = i - row * k;
  const float v = row < rows_valid ? w[(size_t)row * k + col] : 0.0f;
  *reinterpret_cast<__nv_bfloat16*>(dst + elem_offset(rows_total, row, col)) = __float2bfloat16_rn(v);
}
__global__ void k_pack_bias(const float* __restrict__ b1, const float* __restrict__ b2, const float* __restrict__ b3,
                            const float* __restrict__ b_mu, const float* __restrict__ b_value, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < D1) dst[i] = b1[i];
  else if (i < D1 + D2) dst[i] = b2[i - D1];
  else if (i < D1 + D2 + D3) dst[i] = b3[i - D1 - D2];
  else if (i < D1 + D2 + D3 + 12) dst[i] = b_mu[i - D1 - D2 - D3];
  else if (i == D1 + D2 + D3 + 12) dst[i] = b_value[0];
  else if (i < D1 + D2 + D3 + DH) dst[i] = 0.0f;
}

}  // namespace mlp
}  // namespace lm

using namespace lm;
using namespace lm::mlp;

struct LmMlp {
  int device;
  int sm_count;
  uint8_t* blob;
  bool packed;
};

extern "C" {

LmMlp* lm_mlp_create(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { lm_set_error("lm_mlp_create: %s", cudaGetErrorString(e)); return nullptr; }
  if (prop.major != 10) { lm_set_error("lm_mlp_create: tcgen05 kernel needs sm_100a, device is sm_%d%d", prop.major, prop.minor); return nullptr; }
  cudaSetDevice(device);
  LmMlp* m = new LmMlp();
  m->device = device; m->sm_count = prop.multiProcessorCount; m->packed = false; m->blob = nullptr;
  if (cudaMalloc(&m->blob, BLOB_BYTES) != cudaSuccess) { lm_set_error("lm_mlp_create: cudaMalloc failed"); delete m; return nullptr; }
  cudaMemset(m->blob, 0, BLOB_BYTES);
  e = cudaFuncSetAttribute(k_mlp_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) { lm_set_error("lm_mlp_create: smem attribute: %s", cudaGetErrorString(e)); cudaFree(m->blob); delete m; return nullptr; }
  return m;
}

void lm_mlp_destroy(LmMlp* m) {
  if (!m) return;
  cudaFree(m->blob);
  delete m;
}

int lm_mlp_pack(LmMlp* m, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3,
                const float* w_mu, const float* b_mu, const float* w_value, const float* b_value, void* stream) {
  if (!m || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !w_mu || !b_mu || !w_value || !b_value)
    return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_pack: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(m->blob, 0, BLOB_BYTES, s);
  k_pack_operand<<<(D1 * D0 + 255) / 256, 256, 0, s>>>(w1, m->blob + W1_OFF, D1, D1, D0);
  k_pack_operand<<<(D2 * D1 + 255) / 256, 256, 0, s>>>(w2, m->blob + W2_OFF, D2, D2, D1);
  k_pack_operand<<<(D3 * D2 + 255) / 256, 256, 0, s>>>(w3, m->blob + W3_OFF, D3, D3, D2);
  // head rows 0..11 = mean weights, row 12 = value weights (two passes into the same operand)
  k_pack_operand<<<(12 * D3 + 255) / 256, 256, 0, s>>>(w_mu, m->blob + WH_OFF, DH, 12, D3);
  k_pack_operand<<<(1 * D3 + 255) / 256, 256, 0, s>>>(w_value, m->blob + WH_OFF + 12 * 16, DH, 1, D3);
  k_pack_bias<<<2, 256, 0, s>>>(b1, b2, b3, b_mu, b_value, reinterpret_cast<float*>(m->blob + BIAS_OFF));
  m->packed = true;
  return lm_check_launch("lm_mlp_pack");
}

int lm_mlp_forward(LmMlp* m, const float* obs, const float* norm_mean, const float* norm_inv_std, float norm_clip, float* mu,
                   float* value, int32_t n, void* stream) {
  if (!m || !obs || !mu || !value || n <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_forward: bad argument");
  if (!m->packed) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_forward: weights not packed (call lm_mlp_pack)");
  if ((((uintptr_t)obs) | ((uintptr_t)mu)) & 15) return lm_fail(LM_ERR_ALIGNMENT, "lm_mlp_forward: obs/mu must be 16-byte aligned");
  if ((norm_mean == nullptr) != (norm_inv_std == nullptr)) return lm_fail(LM_ERR_BAD_ARG, "lm_mlp_forward: give both norm vectors or neither");
  Args a;
  a.blob = m->blob; a.obs = obs; a.norm_mean = norm_mean; a.norm_inv_std = norm_inv_std; a.norm_clip = norm_clip;
  a.mu = mu; a.value = value; a.n = n; a.num_tiles = (n + TM - 1) / TM;
  const int grid = a.num_tiles < m->sm_count ? a.num_tiles : m->sm_count;
  k_mlp_forward<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(a);
  return lm_check_launch("lm_mlp_forward");
}

}  // extern "C"
