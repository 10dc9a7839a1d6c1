// GNN limb-graph policy forward, fp32 (scripts/graph_model_orebot_ov.py:11-241, hidden 32; wired at
// scripts/skrl_ppo_locomanipulation.py:44-49).
//
// Graph: node 0 = object (obs[0:16]); node 1+j = joint pi(j), pi = (0,1,2,3, 4,6,8,10, 5,7,9,11), features
// (jp, jv, a_cur, a_last) (:117-126).  24 directed edges 0<->{1..4}, k<->k+4 for k in 1..8 (:142-159).
// Layer (:53-59,46): m_ij = ELU(W2 ELU(W1 [h_i ; h_j] + b1) + b2), h_i' = max_j m_ij (i = target).
// W1 [h_i ; h_j] is split per endpoint: P_i = W1a h_i + b1, Q_j = W1b h_j, so each node is
// transformed once instead of once per edge (50 instead of 72 32x32 mat-vecs per layer).
//
// Mapping: TPE = 8 threads per env, each owning CH = 4 of the 32 hidden channels; 40 envs (10 warps) per CTA.  Node
// features h and Q live in shared memory (adjacency is compile-time: the in-edge lists below are staged as
// constants), all weights are staged in shared memory once per CTA in a bank-conflict-free [k/4][j][q] float4 order.
// ELU is exp(x) - 1 on the negative side, as ATen's CPU kernel evaluates it.
#include "lm_common.cuh"
#include "lm_gnn.h"
#include "lm_host.h"

namespace lm {
namespace gnn {

constexpr int H = 32;
constexpr int NODES = 13;
constexpr int ENVS = 40;                  // envs per CTA
constexpr int TPE = 8;                    // threads per env
constexpr int CH = H / TPE;               // hidden channels per thread
constexpr int THREADS = ENVS * TPE;
constexpr int HSTRIDE = NODES * H + 4;    // floats per env in the h / Q buffers (+4: rotate banks between envs)
constexpr int ZSTRIDE = 4 * H + 4;        // up to 4 staged message inputs per env
constexpr int OBS_STRIDE = 64 + 4;

// packed weights (floats).  Matrices are stored as float4 [k/4][j][q] = W[CH q + j][4(k/4) .. +3].
constexpr int W_IN1 = 0;                          // [32 x 16]
constexpr int W_IN2 = W_IN1 + H * 16;             // [32 x 4]
constexpr int W_L = W_IN2 + H * 4;                // per layer: W1a [32x32], W1b [32x32], W2 [32x32]
constexpr int W_LSTRIDE = 3 * H * H;
constexpr int B_IN1 = W_L + 3 * W_LSTRIDE;        // biases: in1, in2, then per layer b1, b2
constexpr int B_IN2 = B_IN1 + H;
constexpr int B_L = B_IN2 + H;
constexpr int W_ACT = B_L + 3 * 2 * H;            // action head W[32], value head W[32], b_act, b_val
constexpr int W_VAL = W_ACT + H;
constexpr int B_HEAD = W_VAL + H;
constexpr int PACKED_FLOATS = ((B_HEAD + 2 + 3) / 4) * 4;
constexpr int SMEM_FLOATS = PACKED_FLOATS + 2 * ENVS * HSTRIDE + ENVS * ZSTRIDE;
constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");
static_assert(ENVS * OBS_STRIDE <= ENVS * HSTRIDE, "obs staging reuses the Q buffer");

__device__ __forceinline__ float elu(float x) { return x > 0.0f ? x : expf(x) - 1.0f; }

// acc[v][j] += sum_k W[CH q+j][k] * x_v[k]  for NV input vectors sharing one weight pass
template <int NV, int K>
__device__ __forceinline__ void matvec8(const float* __restrict__ wp, int q, const float* const x[NV], float acc[NV][CH]) {
  const float4* w4 = reinterpret_cast<const float4*>(wp);
#pragma unroll
  for (int kq = 0; kq < K / 4; ++kq) {
    float4 w[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) w[j] = w4[(kq * CH + j) * TPE + q];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 xv = *reinterpret_cast<const float4*>(x[v] + 4 * kq);
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        acc[v][j] = fmaf(w[j].x, xv.x, acc[v][j]);
        acc[v][j] = fmaf(w[j].y, xv.y, acc[v][j]);
        acc[v][j] = fmaf(w[j].z, xv.z, acc[v][j]);
        acc[v][j] = fmaf(w[j].w, xv.w, acc[v][j]);
      }
    }
  }
}

static_assert(CH == 4, "store8 / load8 move one float4 per thread");
__device__ __forceinline__ void store8(float* dst, const float v[CH]) {
  reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void load8(const float* src, float v[CH]) {
  const float4 a = reinterpret_cast<const float4*>(src)[0];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}

// messages into up to two target nodes handled in one W2 pass: z_e = ELU(P_target + Q_source), NV = total in-edges (<= 4)
template <int NV>
__device__ __forceinline__ void message_pass(const float* __restrict__ w2, const float* __restrict__ b2, int q, float* zenv,
                                             const float* Qenv, const float P[][CH], const int tgt_of[NV], const int src[NV],
                                             float out[NV][CH]) {
  // stage z (each thread its 8 channels), then every thread needs all 32 channels -> warp-level sync (4 lanes of an env share a warp)
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    float qv[CH], z[CH];
    load8(Qenv + src[v] * H + CH * q, qv);
#pragma unroll
    for (int j = 0; j < CH; ++j) z[j] = elu(P[tgt_of[v]][j] + qv[j]);
    store8(zenv + v * H + CH * q, z);
  }
  __syncwarp();
  const float* x[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    x[v] = zenv + v * H;
#pragma unroll
    for (int j = 0; j < CH; ++j) out[v][j] = b2[CH * q + j];
  }
  matvec8<NV, H>(w2, q, x, out);
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < CH; ++j) out[v][j] = elu(out[v][j]);
  __syncwarp();   // z may be overwritten by the next pass
}

struct Args {
  const float* packed;
  const float* obs;
  float* mu;
  float* value;
  int n;
};

__global__ void __launch_bounds__(THREADS, 1) k_gnn_forward(const Args a) {
  extern __shared__ __align__(16) float sm[];
  float* W = sm;
  float* hbuf = sm + PACKED_FLOATS;
  float* qbuf = hbuf + ENVS * HSTRIDE;
  float* zbuf = qbuf + ENVS * HSTRIDE;
  const int tid = threadIdx.x;
  const int le = tid / TPE, q = tid % TPE;       // local env, channel slice
  const int env0 = blockIdx.x * ENVS;
  const int nenv = min(ENVS, a.n - env0);

  // ---- stage weights + the obs tile (obs staged in the Q buffer, unused until layer 1)
  for (int i = tid; i < PACKED_FLOATS / 4; i += THREADS)
    reinterpret_cast<float4*>(W)[i] = __ldg(reinterpret_cast<const float4*>(a.packed) + i);
  for (int i = tid; i < ENVS * 16; i += THREADS) {
    const int e = i >> 4, c4 = i & 15;
    *reinterpret_cast<float4*>(qbuf + e * OBS_STRIDE + 4 * c4) =
        e < nenv ? ld_nc_f4(a.obs + (size_t)(env0 + e) * 64 + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const bool active = le < nenv;
  float* henv = hbuf + le * HSTRIDE;
  float* Qenv = qbuf + le * HSTRIDE;
  float* zenv = zbuf + le * ZSTRIDE;
  const float* obs = qbuf + le * OBS_STRIDE;

  {
    // ---- input layers (:97-104,112-126).  Lanes of padding envs (tail CTA) run the same code on their own
    // smem slots so that the warp-level syncs below stay convergent; only the global stores are guarded.
    {
      float acc[1][CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) acc[0][j] = W[B_IN1 + CH * q + j];
      const float* x[1] = {obs};
      matvec8<1, 16>(W + W_IN1, q, x, acc);
      store8(henv + CH * q, acc[0]);
    }
    // joint nodes: feature k of node 1+j is obs[16 + 12 k + pi(j)]
    const float4* w4 = reinterpret_cast<const float4*>(W + W_IN2);
    float4 w[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) w[j] = w4[j * TPE + q];
#pragma unroll
    for (int jn = 0; jn < 12; ++jn) {
      const int pi = jn < 4 ? jn : (jn < 8 ? 4 + 2 * (jn - 4) : 5 + 2 * (jn - 8));
      const float f0 = obs[16 + pi], f1 = obs[28 + pi], f2 = obs[40 + pi], f3 = obs[52 + pi];
      float y[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        float s = W[B_IN2 + CH * q + j];
        s = fmaf(w[j].x, f0, s); s = fmaf(w[j].y, f1, s); s = fmaf(w[j].z, f2, s); s = fmaf(w[j].w, f3, s);
        y[j] = s;
      }
      store8(henv + (1 + jn) * H + CH * q, y);
    }
  }
  __syncthreads();   // obs staging (Q buffer) fully consumed by every thread of the CTA before Q is written

  {
#pragma unroll 1
    for (int layer = 0; layer < 3; ++layer) {
      const float* w1a = W + W_L + layer * W_LSTRIDE;
      const float* w1b = w1a + H * H;
      const float* w2 = w1b + H * H;
      const float* b1 = W + B_L + layer * 2 * H;
      const float* b2 = b1 + H;
      __syncwarp();
      // ---- Q_n = W1b h_n for all 13 nodes (old h)
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        float acc[4][CH];
        const float* x[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          x[v] = henv + (4 * g + v) * H;
#pragma unroll
          for (int j = 0; j < CH; ++j) acc[v][j] = 0.0f;
        }
        matvec8<4, H>(w1b, q, x, acc);
#pragma unroll
        for (int v = 0; v < 4; ++v) store8(Qenv + (4 * g + v) * H + CH * q, acc[v]);
      }
      {
        float acc[1][CH];
        const float* x[1] = {henv + 12 * H};
#pragma unroll
        for (int j = 0; j < CH; ++j) acc[0][j] = 0.0f;
        matvec8<1, H>(w1b, q, x, acc);
        store8(Qenv + 12 * H + CH * q, acc[0]);
      }
      // ---- P_i = W1a h_i + b1 for all targets, kept in registers level by level; h' written in place
      float Pn[NODES][CH];
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const float* x[4];
        float acc[4][CH];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          x[v] = henv + (4 * g + v) * H;
#pragma unroll
          for (int j = 0; j < CH; ++j) acc[v][j] = b1[CH * q + j];
        }
        matvec8<4, H>(w1a, q, x, acc);
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
          for (int j = 0; j < CH; ++j) Pn[4 * g + v][j] = acc[v][j];
      }
      {
        const float* x[1] = {henv + 12 * H};
        float acc[1][CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) acc[0][j] = b1[CH * q + j];
        matvec8<1, H>(w1a, q, x, acc);
#pragma unroll
        for (int j = 0; j < CH; ++j) Pn[12][j] = acc[0][j];
      }
      __syncwarp();   // every lane of the env has read old h and written its Q slice
      // node 0 <- {1,2,3,4}
      {
        const int tg[4] = {0, 0, 0, 0}, src[4] = {1, 2, 3, 4};
        float out[4][CH], m[CH];
        message_pass<4>(w2, b2, q, zenv, Qenv, Pn, tg, src, out);
#pragma unroll
        for (int j = 0; j < CH; ++j) m[j] = fmaxf(fmaxf(out[0][j], out[1][j]), fmaxf(out[2][j], out[3][j]));
        store8(henv + CH * q, m);
      }
      // nodes 1..4 <- {0, k+4};  nodes 5..8 <- {k-4, k+4}: two targets (4 edges) per W2 pass
#pragma unroll
      for (int pair = 0; pair < 4; ++pair) {
        const int k0 = 1 + 2 * pair, k1 = k0 + 1;
        const int s00 = k0 <= 4 ? 0 : k0 - 4, s10 = k1 <= 4 ? 0 : k1 - 4;
        const int tg[4] = {k0, k0, k1, k1}, src[4] = {s00, k0 + 4, s10, k1 + 4};
        float out[4][CH], m0[CH], m1[CH];
        message_pass<4>(w2, b2, q, zenv, Qenv, Pn, tg, src, out);
#pragma unroll
        for (int j = 0; j < CH; ++j) { m0[j] = fmaxf(out[0][j], out[1][j]); m1[j] = fmaxf(out[2][j], out[3][j]); }
        store8(henv + k0 * H + CH * q, m0);
        store8(henv + k1 * H + CH * q, m1);
      }
      // nodes 9..12 <- {k-4}
      {
        const int tg[4] = {9, 10, 11, 12}, src[4] = {5, 6, 7, 8};
        float out[4][CH];
        message_pass<4>(w2, b2, q, zenv, Qenv, Pn, tg, src, out);
#pragma unroll
        for (int v = 0; v < 4; ++v) store8(henv + (9 + v) * H + CH * q, out[v]);
      }
    }
    __syncwarp();
    // ---- heads: action k = Wa . h[1+k] + ba (:215-226); value = Wv . max_n h[n] + bv (:228-241)
    const int env = env0 + le;
#pragma unroll
    for (int t = 0; t < (12 + TPE - 1) / TPE; ++t) {
      const int k = q + TPE * t;
      if (k < 12) {
        const float* hv = henv + (1 + k) * H;
        float s = W[B_HEAD];
#pragma unroll
        for (int c = 0; c < H; ++c) s = fmaf(W[W_ACT + c], hv[c], s);
        if (active) a.mu[(size_t)env * 12 + k] = s;
      }
    }
    float part = 0.0f;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int c = CH * q + j;
      float m = henv[c];
#pragma unroll
      for (int nidx = 1; nidx < NODES; ++nidx) m = fmaxf(m, henv[nidx * H + c]);
      part = fmaf(W[W_VAL + c], m, part);
    }
#pragma unroll
    for (int off = 1; off < TPE; off <<= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (q == 0 && active) a.value[env] = part + W[B_HEAD + 1];
  }
}

// nn.Linear W [32, ld] columns [col0, col0+K) -> float4 [k/4][j][q]
__global__ void k_pack_matrix(const float* __restrict__ w, float* __restrict__ dst, int ld, int col0, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * K) return;
  const int c = i / K, k = i - c * K;
  const int q = c / CH, j = c % CH;
  dst[(((k >> 2) * CH + j) * TPE + q) * 4 + (k & 3)] = w[(size_t)c * ld + col0 + k];
}
__global__ void k_copy(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

}  // namespace gnn
}  // namespace lm

using namespace lm;
using namespace lm::gnn;

extern "C" {

LmGnn* lm_gnn_create(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { lm_set_error("lm_gnn_create: %s", cudaGetErrorString(e)); return nullptr; }
  if (prop.major != 10) { lm_set_error("lm_gnn_create: library is sm_100a only (device sm_%d%d)", prop.major, prop.minor); return nullptr; }
  cudaSetDevice(device);
  LmGnn* g = new LmGnn();
  g->device = device; g->sm_count = prop.multiProcessorCount; g->ready = false; g->packed = nullptr; g->tc_blob = nullptr; g->train_blob = nullptr; g->timeline = nullptr;
  if (cudaMalloc(&g->packed, PACKED_FLOATS * sizeof(float)) != cudaSuccess) { lm_set_error("lm_gnn_create: cudaMalloc failed"); delete g; return nullptr; }
  cudaMemset(g->packed, 0, PACKED_FLOATS * sizeof(float));
  e = cudaFuncSetAttribute(k_gnn_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) { lm_set_error("lm_gnn_create: smem attribute: %s", cudaGetErrorString(e)); cudaFree(g->packed); delete g; return nullptr; }
  if (lm_gnn_tc_init(g) != LM_OK) { cudaFree(g->packed); delete g; return nullptr; }
  return g;
}

void lm_gnn_destroy(LmGnn* g) {
  if (!g) return;
  cudaFree(g->packed);
  lm_gnn_tc_release(g);
  lm_gnn_train_release(g);
  delete g;
}

// weights, in order: in1.W[32,16] in1.b | in2.W[32,4] in2.b | 3 x (linear1.W[32,64] linear1.b linear2.W[32,32] linear2.b) |
//                    action.W[1,32] action.b[1] | value.W[1,32] value.b[1]            (20 tensors)
int lm_gnn_pack(LmGnn* g, const float* const* w, int32_t num_tensors, void* stream) {
  if (!g || !w || num_tensors != 20) return lm_fail(LM_ERR_BAD_ARG, "lm_gnn_pack: expected 20 tensors");
  for (int i = 0; i < num_tensors; ++i) if (!w[i]) return lm_fail(LM_ERR_BAD_ARG, "lm_gnn_pack: tensor %d is NULL", i);
  cudaStream_t s = (cudaStream_t)stream;
  float* P = g->packed;
  k_pack_matrix<<<(H * 16 + 127) / 128, 128, 0, s>>>(w[0], P + W_IN1, 16, 0, 16);
  k_copy<<<1, 32, 0, s>>>(w[1], P + B_IN1, H);
  k_pack_matrix<<<(H * 4 + 127) / 128, 128, 0, s>>>(w[2], P + W_IN2, 4, 0, 4);
  k_copy<<<1, 32, 0, s>>>(w[3], P + B_IN2, H);
  for (int l = 0; l < 3; ++l) {
    const float* w1 = w[4 + 4 * l];
    k_pack_matrix<<<(H * H + 127) / 128, 128, 0, s>>>(w1, P + W_L + l * W_LSTRIDE, 2 * H, 0, H);            // W1a: columns of h_i
    k_pack_matrix<<<(H * H + 127) / 128, 128, 0, s>>>(w1, P + W_L + l * W_LSTRIDE + H * H, 2 * H, H, H);    // W1b: columns of h_j
    k_copy<<<1, 32, 0, s>>>(w[5 + 4 * l], P + B_L + l * 2 * H, H);
    k_pack_matrix<<<(H * H + 127) / 128, 128, 0, s>>>(w[6 + 4 * l], P + W_L + l * W_LSTRIDE + 2 * H * H, H, 0, H);
    k_copy<<<1, 32, 0, s>>>(w[7 + 4 * l], P + B_L + l * 2 * H + H, H);
  }
  k_copy<<<1, 32, 0, s>>>(w[16], P + W_ACT, H);
  k_copy<<<1, 32, 0, s>>>(w[17], P + B_HEAD, 1);
  k_copy<<<1, 32, 0, s>>>(w[18], P + W_VAL, H);
  k_copy<<<1, 32, 0, s>>>(w[19], P + B_HEAD + 1, 1);
  const int rc = lm_gnn_tc_pack(g, w, stream);          // the same tensors as bf16 tensor-core operands (lm_gnn_forward_bf16)
  if (rc != LM_OK) return rc;
  const int rc2 = lm_gnn_train_pack(g, w, stream);      // and as the fp32 image of the update kernel (lm_gnn_backward)
  if (rc2 != LM_OK) return rc2;
  g->ready = true;
  return lm_check_launch("lm_gnn_pack");
}

int lm_gnn_forward(LmGnn* g, const float* obs, float* mu, float* value, int32_t n, void* stream) {
  if (!g || !obs || !mu || !value || n <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_gnn_forward: bad argument");
  if (!g->ready) return lm_fail(LM_ERR_BAD_ARG, "lm_gnn_forward: weights not packed (call lm_gnn_pack)");
  if (((uintptr_t)obs) & 15) return lm_fail(LM_ERR_ALIGNMENT, "lm_gnn_forward: obs must be 16-byte aligned");
  Args a{g->packed, obs, mu, value, n};
  k_gnn_forward<<<(n + ENVS - 1) / ENVS, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(a);
  return lm_check_launch("lm_gnn_forward");
}

}  // extern "C"
