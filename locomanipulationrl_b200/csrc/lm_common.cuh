// Shared device helpers: contraction-free fp32 arithmetic mirroring torch-CPU eager numerics,
// Isaac quaternion algebra (wxyz), the SciPy float64 quaternion<->matrix path of the reference's
// utils/math.py, and TMA (cp.async.bulk) / mbarrier wrappers for sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lm_b200.h"

namespace lm {

// ---------------------------------------------------------------------------------------------
// fp32 arithmetic with explicit rounding: the reference runs op-by-op torch kernels, so every
// product and sum is rounded separately.  __f*_rn intrinsics are never contracted into FMAs.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }

struct Quat { float w, x, y, z; };
struct Vec3 { float x, y, z; };

// omni.isaac.core.utils.torch.rotations.quat_mul — 8-multiplication Hamilton product
// (call sites QPC:345,430).  Op order restated in oracle/isaac_torch.py.
__device__ __forceinline__ Quat quat_mul(const Quat& a, const Quat& b) {
  float ww = mul(add(a.z, a.x), add(b.x, b.y));
  float yy = mul(sub(a.w, a.y), add(b.w, b.z));
  float zz = mul(add(a.w, a.y), sub(b.w, b.z));
  float xx = add(add(ww, yy), zz);
  float qq = mul(0.5f, add(xx, mul(sub(a.z, a.x), sub(b.x, b.y))));
  Quat r;
  r.w = add(sub(qq, ww), mul(sub(a.z, a.y), sub(b.y, b.z)));
  r.x = add(sub(qq, xx), mul(add(a.x, a.w), add(b.x, b.w)));
  r.y = add(sub(qq, yy), mul(sub(a.w, a.x), add(b.y, b.z)));
  r.z = add(sub(qq, zz), mul(add(a.z, a.y), sub(b.w, b.x)));
  return r;
}

__device__ __forceinline__ Quat quat_conj(const Quat& q) { return Quat{q.w, -q.x, -q.y, -q.z}; }

// torch.cross on CPU evaluates each component as fma(a_i, b_j, -(a_j * b_i)) (measured: the
// vectorised kernel is compiled with FMA contraction; see DESIGN.md "numerics").
__device__ __forceinline__ Vec3 cross_torch(const Vec3& a, const Vec3& b) {
  Vec3 r;
  r.x = fma_rn(a.y, b.z, -mul(a.z, b.y));
  r.y = fma_rn(a.z, b.x, -mul(a.x, b.z));
  r.z = fma_rn(a.x, b.y, -mul(a.y, b.x));
  return r;
}

// quat_rotate / quat_rotate_inverse (QPC:319-329): a = v (2 w^2 - 1); b = cross(u, v) w 2;
// c = u (u . v) 2 with the dot product done by a [1x3]x[3x1] bmm (sequential, no FMA);
// result a + b + c  or  a - b + c.
template <bool kInverse>
__device__ __forceinline__ Vec3 quat_rotate_t(const Quat& q, const Vec3& v) {
  float s = sub(mul(2.0f, mul(q.w, q.w)), 1.0f);
  Vec3 u{q.x, q.y, q.z};
  Vec3 cr = cross_torch(u, v);
  float dot = add(add(mul(u.x, v.x), mul(u.y, v.y)), mul(u.z, v.z));
  Vec3 r;
  float ax = mul(v.x, s), ay = mul(v.y, s), az = mul(v.z, s);
  float bx = mul(mul(cr.x, q.w), 2.0f), by = mul(mul(cr.y, q.w), 2.0f), bz = mul(mul(cr.z, q.w), 2.0f);
  float cx = mul(mul(u.x, dot), 2.0f), cy = mul(mul(u.y, dot), 2.0f), cz = mul(mul(u.z, dot), 2.0f);
  if (kInverse) {
    r.x = add(sub(ax, bx), cx); r.y = add(sub(ay, by), cy); r.z = add(sub(az, bz), cz);
  } else {
    r.x = add(add(ax, bx), cx); r.y = add(add(ay, by), cy); r.z = add(add(az, bz), cz);
  }
  return r;
}

// torch.norm(p=2) on CPU accumulates with fma(x, x, acc) and takes a correctly rounded sqrt.
__device__ __forceinline__ float norm2_torch(float a, float b) { return __fsqrt_rn(fma_rn(b, b, mul(a, a))); }
__device__ __forceinline__ float norm3_torch(float a, float b, float c) {
  return __fsqrt_rn(fma_rn(c, c, fma_rn(b, b, mul(a, a))));
}

// ---------------------------------------------------------------------------------------------
// utils/math.py:7-19 — rot_matrices_from_quat: SciPy Rotation.from_quat (normalise) -> as_matrix
// in float64, cast to float32.  m is row-major 3x3.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rot_from_quat_scipy(const Quat& q, float m[9]) {
  double w = q.w, x = q.x, y = q.y, z = q.z;
  // SciPy divides by the norm; multiplying by rsqrt differs by <= 1 ulp of double, far below the fp32 cast
  const double inv_n = rsqrt(x * x + y * y + z * z + w * w);
  x *= inv_n; y *= inv_n; z *= inv_n; w *= inv_n;
  double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
  double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
  m[0] = (float)(x2 - y2 - z2 + w2); m[1] = (float)(2.0 * (xy - zw)); m[2] = (float)(2.0 * (xz + yw));
  m[3] = (float)(2.0 * (xy + zw)); m[4] = (float)(-x2 + y2 - z2 + w2); m[5] = (float)(2.0 * (yz - xw));
  m[6] = (float)(2.0 * (xz - yw)); m[7] = (float)(2.0 * (yz + xw)); m[8] = (float)(-x2 - y2 + z2 + w2);
}

// torch.bmm of tiny matrices runs the naive CPU kernel: sequential k, separate multiply and add.
// (R v)_i, row i of R.
__device__ __forceinline__ float row_dot(const float* r, int stride, const Vec3& v) {
  return add(add(mul(r[0], v.x), mul(r[stride], v.y)), mul(r[2 * stride], v.z));
}

// utils/math.py:102-113 — transform_vectors: R v + t through the homogeneous 4x4 bmm.
__device__ __forceinline__ Vec3 tf_apply(const float m[9], const Vec3& t, const Vec3& v) {
  return Vec3{add(row_dot(m + 0, 1, v), t.x), add(row_dot(m + 3, 1, v), t.y), add(row_dot(m + 6, 1, v), t.z)};
}
// utils/math.py:115-127,152-173 — inverse transform: R^T v + p', p' = bmm(R^T, -t).
__device__ __forceinline__ Vec3 tf_inverse_translation(const float m[9], const Vec3& t) {
  Vec3 nt{-t.x, -t.y, -t.z};
  return Vec3{row_dot(m + 0, 3, nt), row_dot(m + 1, 3, nt), row_dot(m + 2, 3, nt)};
}
__device__ __forceinline__ Vec3 tf_inverse_apply(const float m[9], const Vec3& pinv, const Vec3& v) {
  return Vec3{add(row_dot(m + 0, 3, v), pinv.x), add(row_dot(m + 1, 3, v), pinv.y), add(row_dot(m + 2, 3, v), pinv.z)};
}
__device__ __forceinline__ float tf_inverse_apply_z(const float m[9], float pinv_z, const Vec3& v) {
  return add(row_dot(m + 2, 3, v), pinv_z);
}

// c = a(^T) b for 3x3 fp32, naive bmm order.
template <bool kTransposeA>
__device__ __forceinline__ void mat3_mul(const float a[9], const float b[9], float c[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a0 = kTransposeA ? a[0 + i] : a[3 * i + 0];
      float a1 = kTransposeA ? a[3 + i] : a[3 * i + 1];
      float a2 = kTransposeA ? a[6 + i] : a[3 * i + 2];
      c[3 * i + j] = add(add(mul(a0, b[j]), mul(a1, b[3 + j])), mul(a2, b[6 + j]));
    }
}

// utils/math.py:21-31 — quat_from_rot_matrices: SciPy Rotation.from_matrix in float64.  SciPy
// (1.18, cython backend) replaces a not-exactly-orthogonal input by its polar factor U V^T (SVD),
// then applies the largest-of-(diag, trace) branch and normalises.  The polar factor is obtained
// here by ONE Newton step X <- (X + X^-T)/2: the input is a product of two fp32 rotation matrices,
// orthogonal to e ~ 1e-7, and the step leaves e^2/2 ~ 5e-15 — 7 orders below the fp32 rounding of
// the result (a second step, used at first, changed no fp32 output in any test).  Returns wxyz with
// w >= 0 (math.py:52-54,78-80).
__device__ __forceinline__ Quat quat_from_rot_scipy(const float m32[9]) {
  double a[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) a[i] = (double)m32[i];
  {
    double c[9];
    c[0] = a[4] * a[8] - a[5] * a[7]; c[1] = a[5] * a[6] - a[3] * a[8]; c[2] = a[3] * a[7] - a[4] * a[6];
    c[3] = a[2] * a[7] - a[1] * a[8]; c[4] = a[0] * a[8] - a[2] * a[6]; c[5] = a[1] * a[6] - a[0] * a[7];
    c[6] = a[1] * a[5] - a[2] * a[4]; c[7] = a[2] * a[3] - a[0] * a[5]; c[8] = a[0] * a[4] - a[1] * a[3];
    const double inv_det = __drcp_rn(a[0] * c[0] + a[1] * c[1] + a[2] * c[2]);
#pragma unroll
    for (int i = 0; i < 9; ++i) a[i] = 0.5 * (a[i] + c[i] * inv_det);
  }
  double tr = a[0] + a[4] + a[8];
  double d0 = a[0], d1 = a[4], d2 = a[8];
  double qx, qy, qz, qw;
  // argmax over (m00, m11, m22, trace), first maximum wins
  int choice = 0; double best = d0;
  if (d1 > best) { best = d1; choice = 1; }
  if (d2 > best) { best = d2; choice = 2; }
  if (tr > best) { best = tr; choice = 3; }
  if (choice == 0) {
    qx = 1.0 - tr + 2.0 * a[0]; qy = a[3] + a[1]; qz = a[6] + a[2]; qw = a[7] - a[5];
  } else if (choice == 1) {
    qy = 1.0 - tr + 2.0 * a[4]; qz = a[7] + a[5]; qx = a[1] + a[3]; qw = a[2] - a[6];
  } else if (choice == 2) {
    qz = 1.0 - tr + 2.0 * a[8]; qx = a[2] + a[6]; qy = a[5] + a[7]; qw = a[3] - a[1];
  } else {
    qx = a[7] - a[5]; qy = a[2] - a[6]; qz = a[3] - a[1]; qw = 1.0 + tr;
  }
  const double inv_n = rsqrt(qx * qx + qy * qy + qz * qz + qw * qw);
  Quat q{(float)(qw * inv_n), (float)(qx * inv_n), (float)(qy * inv_n), (float)(qz * inv_n)};
  if (q.w < 0.0f) { q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  return q;
}

// quat_from_euler_xyz (Isaac) with the reference's angle draw (utils/math.py:187-191).  sinf/cosf (<= 1 ulp, like
// torch-CPU's SLEEF kernels; the two libms agree bit-for-bit in ~95 % of draws and differ by 1 ulp otherwise).
__device__ __forceinline__ Quat goal_from_uniforms(const LmTaskParams& p, float u0, float u1, float u2) {
  float roll = add(p.min_roll, mul(p.roll_range, u0));
  float pitch = add(p.min_pitch, mul(p.pitch_range, u1));
  float yaw = add(p.min_yaw, mul(p.yaw_range, u2));
  float sy, cy, sr, cr, sp, cp;
  sincosf(mul(yaw, 0.5f), &sy, &cy);
  sincosf(mul(roll, 0.5f), &sr, &cr);
  sincosf(mul(pitch, 0.5f), &sp, &cp);
  Quat q;
  q.w = add(mul(mul(cy, cr), cp), mul(mul(sy, sr), sp));
  q.x = sub(mul(mul(cy, sr), cp), mul(mul(sy, cr), sp));
  q.y = add(mul(mul(cy, cr), sp), mul(mul(sy, sr), cp));
  q.z = sub(mul(mul(sy, cr), cp), mul(mul(cy, sr), sp));
  return q;
}

// ---------------------------------------------------------------------------------------------
// TMA bulk copies + mbarrier (sm_90+/sm_100a).  Every per-env array is row-major [N, k], so the
// rows of one env tile are ONE contiguous span: a 1-D cp.async.bulk moves it without a tensor map.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a mis-programmed transaction count traps (sticky error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global (bulk async group)
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy smem writes visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start while
// its predecessor on the stream is still draining; pdl_wait() blocks until that predecessor has completed and its writes
// are visible, so everything before it (barrier init, TMEM allocation, resident-weight TMA) overlaps the predecessor's
// tail.  pdl_launch_dependents() lets the successor be scheduled as soon as every CTA of this grid has started.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float4 ld_nc_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

}  // namespace lm
