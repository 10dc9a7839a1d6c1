// tcgen05 / TMEM wrappers and the UMMA canonical (no-swizzle, K-major) operand layout shared by the tensor-core
// kernels (lm_mlp.cu, lm_gnn_tc.cu).
#pragma once
#include <cuda_bf16.h>

#include "lm_common.cuh"

namespace lm {
namespace tc {

// UMMA canonical K-major, no swizzle: core matrix = 8 rows x 16 B, rows 16 B apart; core matrices
// 128 B apart along M/N (SBO); K chunks (8 bf16) `lbo` bytes apart.  lbo = rows*16 + 16: the extra
// 16 B rotates banks so writing 8 chunks of one row is conflict free.
__host__ __device__ constexpr int lbo_bytes(int rows) { return rows * 16 + 16; }
__host__ __device__ constexpr int operand_bytes(int rows, int k) { return (k / 8) * lbo_bytes(rows); }
__host__ __device__ constexpr int elem_offset(int rows, int row, int k) {   // byte offset of element (row, k)
  return (k / 8) * lbo_bytes(rows) + row * 16 + (k % 8) * 2;
}

// One lane of a converged warp (elect.sync).  Code under `if (elect_one())` is known by ptxas to run on a single thread, so
// tcgen05.mma operands stay in uniform registers; under `if (lane == 0)` it wraps every MMA in a per-lane ELECT/BRA loop
// (measured: 62 cycles per issued MMA in the control thread, independent of the MMA shape).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xFFFFFFFF;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred));
  return pred != 0;
}

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
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t r[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t r[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
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

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// nn.ELU(alpha=1) = max(x, min(exp(x) - 1, 0)): exp(x) - 1 >= x everywhere, and it is > 0 exactly when x > 0
__device__ __forceinline__ float elu(float x) {
  return fmaxf(x, fminf(ex2_approx(x * 1.4426950408889634f) - 1.0f, 0.0f));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 — one issue slot for two lanes' worth of work)
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// two accumulator columns -> +bias, ELU -> packed bf16x2.  ELU(x) = max(x, min(exp(x) - 1, 0)).
__device__ __forceinline__ uint32_t elu_pair(uint32_t acc_lo, uint32_t acc_hi, float b_lo, float b_hi) {
  const uint64_t x2 = add2(pk2u(acc_lo, acc_hi), pk2(b_lo, b_hi));
  float t0, t1, x0, x1, r0, r1;
  upk2(mul2(x2, pk2(1.4426950408889634f, 1.4426950408889634f)), t0, t1);
  upk2(add2(pk2(ex2_approx(t0), ex2_approx(t1)), pk2(-1.0f, -1.0f)), r0, r1);
  upk2(x2, x0, x1);
  return pack_bf16(fmaxf(x0, fminf(r0, 0.0f)), fmaxf(x1, fminf(r1, 0.0f)));
}
// The same pair in 8 instructions instead of 10: ELU(x) = min(e^x - 1, relu(x)) — for x > 0, e^x - 1 >= x; for x <= 0, e^x - 1 <= 0 —
// with relu folded into the conversion (cvt.rn.relu.bf16x2) and the min taken on the packed bf16 pair (rounding is monotone, so
// min(bf16(a), bf16(b)) = bf16(min(a, b))).  No FMNMX on the half-rate ALU pipe, one packed add for both "- 1".  Differs from elu_pair
// only where ex2.approx returns a value below 1 for a tiny positive argument (|difference| <= 2^-22).
__device__ __forceinline__ uint32_t elu_pair_relu(uint32_t acc_lo, uint32_t acc_hi, float b_lo, float b_hi) {
  const uint64_t x2 = add2(pk2u(acc_lo, acc_hi), pk2(b_lo, b_hi));
  float t0, t1, x0, x1, r0, r1;
  upk2(mul2(x2, pk2(1.4426950408889634f, 1.4426950408889634f)), t0, t1);
  upk2(x2, x0, x1);
  upk2(add2(pk2(ex2_approx(t0), ex2_approx(t1)), pk2(-1.0f, -1.0f)), r0, r1);
  uint32_t pos, neg, out;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pos) : "f"(x1), "f"(x0));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(neg) : "f"(r1), "f"(r0));
  asm("min.bf16x2 %0, %1, %2;" : "=r"(out) : "r"(pos), "r"(neg));
  return out;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c);
// The pair in 7 instructions, for activations stored SCALED by s = log2(e): with h'' = s h as the A operand, the accumulator of the
// unscaled weights is s (W h), so t = acc + s b IS the ex2 argument (no multiply), and the scaled output is
// s ELU(t / s) = min(s 2^t - s, relu(t)): FADD2, 2 MUFU, FFMA2, 2 F2FP (one with relu), HMNMX2.
__device__ __forceinline__ uint32_t elu_pair_scaled(uint32_t acc_lo, uint32_t acc_hi, float sb_lo, float sb_hi) {
  constexpr float kS = 1.4426950408889634f;
  const uint64_t t2 = add2(pk2u(acc_lo, acc_hi), pk2(sb_lo, sb_hi));
  float t0, t1, r0, r1;
  upk2(t2, t0, t1);
  upk2(fma2(pk2(ex2_approx(t0), ex2_approx(t1)), pk2(kS, kS), pk2(-kS, -kS)), r0, r1);
  uint32_t pos, neg, out;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pos) : "f"(t1), "f"(t0));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(neg) : "f"(r1), "f"(r0));
  asm("min.bf16x2 %0, %1, %2;" : "=r"(out) : "r"(pos), "r"(neg));
  return out;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// The same ELU pair with exp on the FMA pipe instead of MUFU (the epilogues are MUFU-bound: 16 ex2/clk/SM, so a share of
// the pairs is moved to the otherwise idle FMA pipe).  t = x log2(e) is clamped to [-64, 0] by ONE saturating FMA
// (s = sat(t/64 + 1)); n = round(64 s) appears in the low mantissa bits of 64 s + (1.5 * 2^23 - 64); 2^f, f = 64 s - n in
// [-0.5, 0.5], is a cubic with p(0) = 1 (max relative error 1.0e-4, 20x below the bf16 rounding of the result) and the
// exponent is added as an integer.  exp <= 1 by construction, so ELU = max(x, exp - 1) needs no min().
__device__ __forceinline__ uint32_t elu_pair_fma(uint32_t acc_lo, uint32_t acc_hi, float b_lo, float b_hi) {
  constexpr float kMagic = 12582912.0f - 64.0f;
  const uint64_t x2 = add2(pk2u(acc_lo, acc_hi), pk2(b_lo, b_hi));
  float x0, x1;
  upk2(x2, x0, x1);
  const uint64_t s2 = pk2(__saturatef(fmaf(x0, 1.4426950408889634f / 64.0f, 1.0f)), __saturatef(fmaf(x1, 1.4426950408889634f / 64.0f, 1.0f)));
  const uint64_t k64 = pk2(64.0f, 64.0f);
  const uint64_t r2 = fma2(s2, k64, pk2(kMagic, kMagic));
  const uint64_t f2 = fma2(s2, k64, fma2(r2, pk2(-1.0f, -1.0f), pk2(kMagic, kMagic)));     // 64 s - n
  uint64_t p2 = fma2(f2, pk2(0.05500871f, 0.05500871f), pk2(0.24221068f, 0.24221068f));
  p2 = fma2(p2, f2, pk2(0.69328292f, 0.69328292f));
  p2 = fma2(p2, f2, pk2(1.0f, 1.0f));
  float r0, r1, p0, p1;
  upk2(r2, r0, r1);
  upk2(p2, p0, p1);
  const float e0 = __uint_as_float((__float_as_uint(r0) << 23) + __float_as_uint(p0));
  const float e1 = __uint_as_float((__float_as_uint(r1) << 23) + __float_as_uint(p1));
  float y0, y1;
  upk2(add2(pk2(e0, e1), pk2(-1.0f, -1.0f)), y0, y1);
  return pack_bf16(fmaxf(x0, y0), fmaxf(x1, y1));
}
// elu_pair_fma for activations stored scaled by log2(e) (see elu_pair_scaled): t = acc + s b is the exponent itself
__device__ __forceinline__ uint32_t elu_pair_fma_scaled(uint32_t acc_lo, uint32_t acc_hi, float sb_lo, float sb_hi) {
  constexpr float kMagic = 12582912.0f - 64.0f, kS = 1.4426950408889634f;
  const uint64_t t2 = add2(pk2u(acc_lo, acc_hi), pk2(sb_lo, sb_hi));
  float t0, t1;
  upk2(t2, t0, t1);
  const uint64_t s2 = pk2(__saturatef(fmaf(t0, 1.0f / 64.0f, 1.0f)), __saturatef(fmaf(t1, 1.0f / 64.0f, 1.0f)));
  const uint64_t k64 = pk2(64.0f, 64.0f);
  const uint64_t r2 = fma2(s2, k64, pk2(kMagic, kMagic));
  const uint64_t f2 = fma2(s2, k64, fma2(r2, pk2(-1.0f, -1.0f), pk2(kMagic, kMagic)));     // 64 s - n
  uint64_t p2 = fma2(f2, pk2(0.05500871f, 0.05500871f), pk2(0.24221068f, 0.24221068f));
  p2 = fma2(p2, f2, pk2(0.69328292f, 0.69328292f));
  p2 = fma2(p2, f2, pk2(1.0f, 1.0f));
  float r0, r1, p0, p1;
  upk2(r2, r0, r1);
  upk2(p2, p0, p1);
  const float e0 = __uint_as_float((__float_as_uint(r0) << 23) + __float_as_uint(p0));
  const float e1 = __uint_as_float((__float_as_uint(r1) << 23) + __float_as_uint(p1));
  float y0, y1;
  upk2(fma2(pk2(e0, e1), pk2(kS, kS), pk2(-kS, -kS)), y0, y1);
  return pack_bf16(fmaxf(t0, y0), fmaxf(t1, y1));
}
template <int kFma>
__device__ __forceinline__ uint32_t elu_pair_mix_scaled(int i, uint32_t acc_lo, uint32_t acc_hi, float sb_lo, float sb_hi) {
  return i < kFma ? elu_pair_fma_scaled(acc_lo, acc_hi, sb_lo, sb_hi) : elu_pair_scaled(acc_lo, acc_hi, sb_lo, sb_hi);
}
// pair `i` of a group of four: the first kFma pairs go to the FMA pipe
template <int kFma>
__device__ __forceinline__ uint32_t elu_pair_mix(int i, uint32_t acc_lo, uint32_t acc_hi, float b_lo, float b_hi) {
#ifdef LM_DIAG_NO_EX2             // diagnostic (wrong results): no exponential at all -> is the kernel MUFU / issue bound?
  return pack_bf16(__uint_as_float(acc_lo) + b_lo, __uint_as_float(acc_hi) + b_hi);
#endif
#ifdef LM_ELU_MINMAX            // the FMNMX form (measured 4 % slower in k_mlp_forward: 4 half-rate ALU instructions per pair)
  return i < kFma ? elu_pair_fma(acc_lo, acc_hi, b_lo, b_hi) : elu_pair(acc_lo, acc_hi, b_lo, b_hi);
#else
  return i < kFma ? elu_pair_fma(acc_lo, acc_hi, b_lo, b_hi) : elu_pair_relu(acc_lo, acc_hi, b_lo, b_hi);
#endif
}

}  // namespace tc
}  // namespace lm
