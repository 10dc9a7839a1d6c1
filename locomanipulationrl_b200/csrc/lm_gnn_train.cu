// PPO update of the GNN limb-graph policy: gradients of (mu, value) w.r.t. every parameter in one kernel.
//
// Network: scripts/graph_model_orebot_ov.py:11-241 (hidden 32), restated in locomanipulationrl_b200/policy/gnn.py::reference_forward
// (whose torch autograd is the checker of this kernel).  Per env:
//   h0[0] = in1(obs[0:16]); h0[1+j] = in2([jp, jv, a_cur, a_last] of joint pi(j))
//   3 x { P = W1a h + b1; Q = W1b h (per node);  per edge (s -> t): z1 = ELU(P_t + Q_s); m = ELU(W2 z1 + b2);  h'_t = max over in-edges m }
//   mu_k = w_act . h3[1+k] + b_act;  value = w_val . max_n h3[n] + b_val
// k_gnn_backward: one WARP per env (grid-stride), lane = hidden channel.  Every 32-vector lives as one register per lane; a
// mat-vec batch y_v = W x_v shares each weight fetch (one conflict-free 128-bit shared-memory load per four k) over all vectors
// of the batch, the x_v are read back from a per-warp scratch as broadcast 128-bit loads.  The layer activations are NOT stored
// by the forward pass (23 KB per env): the kernel recomputes the forward in fp32 (keeping the three layer inputs in its scratch),
// then walks the layers backwards.  Weight gradients: lane c first sums its row's contributions over the nodes / edges of the env
// in registers, then adds ONE value per (k, c) into the CTA's gradient image in shared memory (bank = lane: conflict-free);
// the images leave as per-CTA partials and k_gnn_grad_reduce sums them in CTA order into the nn.Linear layouts.
// The max-aggregation routes the gradient to the first in-edge attaining the maximum (ties have measure zero).
#include <string.h>

#include "lm_common.cuh"
#include "lm_gnn.h"
#include "lm_host.h"

namespace lm {
namespace gnn_train {

constexpr int H = 32, NODES = 13, EDGES = 24;
constexpr int WARPS = 8, THREADS = WARPS * 32;

// edges (s -> t) grouped by target: t = 0: s = 1..4 | t = 1..4: s = 0, t + 4 | t = 5..8: s = t - 4, t + 4 | t = 9..12: s = t - 4
__host__ __device__ constexpr int etgt(int e) { return e < 4 ? 0 : e < 12 ? 1 + (e - 4) / 2 : e < 20 ? 5 + (e - 12) / 2 : 9 + (e - 20); }
__host__ __device__ constexpr int esrc(int e) {
  return e < 4 ? 1 + e : e < 12 ? (((e - 4) & 1) ? etgt(e) + 4 : 0) : e < 20 ? (((e - 12) & 1) ? etgt(e) + 4 : etgt(e) - 4) : etgt(e) - 4;
}
__host__ __device__ constexpr int efirst(int t) { return t == 0 ? 0 : t < 5 ? 4 + 2 * (t - 1) : t < 9 ? 12 + 2 * (t - 5) : 20 + (t - 9); }
__host__ __device__ constexpr int ecount(int t) { return t == 0 ? 4 : t < 9 ? 2 : 1; }
// joint of node 1 + j (graph_model_orebot_ov.py:117-126)
__host__ __device__ constexpr int joint_of(int j) { return j < 4 ? j : j < 8 ? 4 + 2 * (j - 4) : 5 + 2 * (j - 8); }

// weight image (floats): matrices in the "packed" form [k/4][32 lanes][4]: lane c reads W[c][4q..4q+3] as one float4, conflict free.
// `P` = forward (lane = output channel), `T` = transposed (lane = input channel: y[k] = sum_c W[c][k] d[c]).
constexpr int WP_IN1 = 0;                       // [32 x 16]
constexpr int WP_IN2 = WP_IN1 + H * 16;         // [32 x 4]
constexpr int WL = WP_IN2 + H * 4;              // per layer: W1aP, W1bP, W2P, W1aT, W1bT, W2T
constexpr int WL_STRIDE = 6 * H * H;
constexpr int B_IN1 = WL + 3 * WL_STRIDE, B_IN2 = B_IN1 + H, B_L = B_IN2 + H;      // per layer b1, b2
constexpr int W_ACT = B_L + 3 * 2 * H, W_VAL = W_ACT + H, B_HEAD = W_VAL + H;      // b_act, b_val
constexpr int BLOB_FLOATS = ((B_HEAD + 2 + 3) / 4) * 4;

// gradient image (floats): matrices stored [k][32] (k-major: lane c adds at k * 32 + c)
constexpr int G_IN1 = 0, GB_IN1 = G_IN1 + 16 * H, G_IN2 = GB_IN1 + H, GB_IN2 = G_IN2 + 4 * H, GL = GB_IN2 + H;
constexpr int GL_W1 = 0, GL_B1 = 64 * H, GL_W2 = GL_B1 + H, GL_B2 = GL_W2 + H * H, GL_STRIDE = GL_B2 + H;
constexpr int G_ACT = GL + 3 * GL_STRIDE, GB_ACT = G_ACT + H, G_VAL = GB_ACT + 1, GB_VAL = G_VAL + H;
constexpr int G_FLOATS = GB_VAL + 1;            // 10,178 = the parameter count
constexpr int G_PAD = ((G_FLOATS + 3) / 4) * 4;

// per-warp scratch (floats): the obs row, the layer input as broadcast source, 26 staged vectors, the three saved layer inputs
constexpr int S_OBS = 0, S_H = 64, S_V = S_H + NODES * H, S_SAVE = S_V + 26 * H, S_WARP = S_SAVE + 3 * NODES * H;
constexpr int SMEM_FLOATS = BLOB_FLOATS + G_PAD + WARPS * S_WARP;
constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
static_assert(SMEM_BYTES <= 232448 - 1024, "exceeds 227 KB");
static_assert(S_WARP % 4 == 0 && S_V % 4 == 0 && S_SAVE % 4 == 0 && BLOB_FLOATS % 4 == 0 && G_PAD % 4 == 0, "float4 alignment");

__device__ __forceinline__ float elu(float x) { return x > 0.0f ? x : expf(x) - 1.0f; }
__device__ __forceinline__ float elu_grad_from_out(float y) { return y > 0.0f ? 1.0f : y + 1.0f; }    // ELU'(x) given y = ELU(x)

// acc[v] += sum_k W[lane][k] * X[v][k]   (W packed, K = 32; X: NV consecutive 32-vectors in shared memory)
template <int NV>
__device__ __forceinline__ void mv_batch(const float* __restrict__ wp, const float* __restrict__ x, int lane, float (&acc)[NV]) {
  const float4* w4 = reinterpret_cast<const float4*>(wp);
  const float4* x4 = reinterpret_cast<const float4*>(x);
#pragma unroll
  for (int kq = 0; kq < H / 4; ++kq) {
    const float4 w = w4[kq * 32 + lane];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 xv = x4[v * (H / 4) + kq];
      acc[v] = fmaf(w.x, xv.x, fmaf(w.y, xv.y, fmaf(w.z, xv.z, fmaf(w.w, xv.w, acc[v]))));
    }
  }
}

// G[k][lane] += sum_v d[v] * X[v][k]  for k = 0..31   (d: this lane's element of NV gradient vectors; X in shared memory)
template <int NV>
__device__ __forceinline__ void outer_acc(float* __restrict__ g, const float (&d)[NV], const float* __restrict__ x, int lane) {
  const float4* x4 = reinterpret_cast<const float4*>(x);
#pragma unroll
  for (int kq = 0; kq < H / 4; ++kq) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 xv = x4[v * (H / 4) + kq];
      s0 = fmaf(d[v], xv.x, s0); s1 = fmaf(d[v], xv.y, s1); s2 = fmaf(d[v], xv.z, s2); s3 = fmaf(d[v], xv.w, s3);
    }
    atomicAdd(g + (4 * kq + 0) * H + lane, s0); atomicAdd(g + (4 * kq + 1) * H + lane, s1);
    atomicAdd(g + (4 * kq + 2) * H + lane, s2); atomicAdd(g + (4 * kq + 3) * H + lane, s3);
  }
}

// one graph layer forward from the layer input staged in `hs` (shared, [13][32]): P, Q per node, messages m per edge, h' per node.
// `sv` receives z1 per edge ([24][32]).
__device__ __forceinline__ void layer_forward(const float* __restrict__ wl, const float* __restrict__ bl, const float* __restrict__ hs, float* __restrict__ sv,
                                              int lane, float (&P)[NODES], float (&Q)[NODES], float (&m)[EDGES], float (&hout)[NODES]) {
  const float b1 = bl[lane], b2 = bl[H + lane];
#pragma unroll
  for (int n = 0; n < NODES; ++n) { P[n] = b1; Q[n] = 0.0f; }
  mv_batch<NODES>(wl + 0 * H * H, hs, lane, P);
  mv_batch<NODES>(wl + 1 * H * H, hs, lane, Q);
#pragma unroll
  for (int e = 0; e < EDGES; ++e) sv[e * H + lane] = elu(P[etgt(e)] + Q[esrc(e)]);
  __syncwarp();
#pragma unroll
  for (int e = 0; e < EDGES; ++e) m[e] = b2;
  mv_batch<EDGES>(wl + 2 * H * H, sv, lane, m);
#pragma unroll
  for (int e = 0; e < EDGES; ++e) m[e] = elu(m[e]);
#pragma unroll
  for (int t = 0; t < NODES; ++t) {
    float best = m[efirst(t)];
#pragma unroll
    for (int i = 1; i < ecount(t); ++i) best = fmaxf(best, m[efirst(t) + i]);
    hout[t] = best;
  }
}

struct Args {
  const float* blob; const float* obs; const float* gmu; const float* gval;
  int n;
  float* partials;      // [grid][G_FLOATS]
};

__global__ void __launch_bounds__(THREADS, 1) k_gnn_backward(const Args a) {
  extern __shared__ __align__(16) float sm[];
  float* W = sm;
  float* G = sm + BLOB_FLOATS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* scr = sm + BLOB_FLOATS + G_PAD + warp * S_WARP;
  float* OBS = scr + S_OBS; float* HS = scr + S_H; float* SV = scr + S_V; float* SAVE = scr + S_SAVE;
  {
    const float4* src = reinterpret_cast<const float4*>(a.blob);
    float4* dst = reinterpret_cast<float4*>(W);
    for (int i = tid; i < BLOB_FLOATS / 4; i += THREADS) dst[i] = __ldg(src + i);
    for (int i = tid; i < G_PAD; i += THREADS) G[i] = 0.0f;
  }
  __syncthreads();

  for (int env = blockIdx.x * WARPS + warp; env < a.n; env += gridDim.x * WARPS) {
    // ---------------------------------------------------------------- forward (fp32), layer inputs saved in the scratch
    __syncwarp();
    OBS[lane] = __ldg(a.obs + (size_t)env * 64 + lane);
    OBS[32 + lane] = __ldg(a.obs + (size_t)env * 64 + 32 + lane);
    __syncwarp();
    float h[NODES];
    {
      const float4* w4 = reinterpret_cast<const float4*>(W + WP_IN1);
      float acc = W[B_IN1 + lane];
#pragma unroll
      for (int kq = 0; kq < 4; ++kq) {
        const float4 w = w4[kq * 32 + lane];
        const float4 x = reinterpret_cast<const float4*>(OBS)[kq];
        acc = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, acc))));
      }
      h[0] = acc;
      const float4 w2 = reinterpret_cast<const float4*>(W + WP_IN2)[lane];
      const float b = W[B_IN2 + lane];
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const int jj = joint_of(j);
        h[1 + j] = fmaf(w2.x, OBS[16 + jj], fmaf(w2.y, OBS[28 + jj], fmaf(w2.z, OBS[40 + jj], fmaf(w2.w, OBS[52 + jj], b))));
      }
    }
    float P[NODES], Q[NODES], m[EDGES];
#pragma unroll 1
    for (int l = 0; l < 3; ++l) {
#pragma unroll
      for (int n = 0; n < NODES; ++n) { HS[n * H + lane] = h[n]; SAVE[(l * NODES + n) * H + lane] = h[n]; }
      __syncwarp();
      float hn[NODES];
      layer_forward(W + WL + l * WL_STRIDE, W + B_L + l * 2 * H, HS, SV, lane, P, Q, m, hn);
      __syncwarp();
#pragma unroll
      for (int n = 0; n < NODES; ++n) h[n] = hn[n];
    }
    // ---------------------------------------------------------------- heads: d(mu, value) -> dh3, head gradients
    float dh[NODES];
    {
      const float w_act = W[W_ACT + lane], w_val = W[W_VAL + lane];
      const float dv = __ldg(a.gval + env);
      float g_act = 0.0f, s_mu = 0.0f;
      dh[0] = 0.0f;
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const float dm = __ldg(a.gmu + (size_t)env * 12 + k);
        dh[1 + k] = dm * w_act;
        g_act = fmaf(dm, h[1 + k], g_act);
        s_mu += dm;
      }
      float best = h[0];
      int arg = 0;
#pragma unroll
      for (int n = 1; n < NODES; ++n) if (h[n] > best) { best = h[n]; arg = n; }
#pragma unroll
      for (int n = 0; n < NODES; ++n) if (arg == n) dh[n] = fmaf(dv, w_val, dh[n]);
      atomicAdd(G + G_ACT + lane, g_act);
      atomicAdd(G + G_VAL + lane, dv * best);
      if (lane == 0) { atomicAdd(G + GB_ACT, s_mu); atomicAdd(G + GB_VAL, dv); }
    }
    // ---------------------------------------------------------------- layers, last to first (activations recomputed from the saved input)
#pragma unroll 1
    for (int l = 2; l >= 0; --l) {
      const float* wl = W + WL + l * WL_STRIDE;
      float* gl = G + GL + l * GL_STRIDE;
      __syncwarp();
#pragma unroll
      for (int n = 0; n < NODES; ++n) HS[n * H + lane] = SAVE[(l * NODES + n) * H + lane];
      __syncwarp();
      float hn[NODES];
      layer_forward(wl, W + B_L + l * 2 * H, HS, SV, lane, P, Q, m, hn);          // SV = z1 per edge, m = ELU(z2) per edge
      // max-aggregation: the gradient of h'_t goes to the first in-edge attaining the maximum, then through ELU'(z2)
      float dz2[EDGES];
      float db2 = 0.0f;
#pragma unroll
      for (int t = 0; t < NODES; ++t) {
        int arg = 0;
        float best = m[efirst(t)];
#pragma unroll
        for (int i = 1; i < ecount(t); ++i) if (m[efirst(t) + i] > best) { best = m[efirst(t) + i]; arg = i; }
#pragma unroll
        for (int i = 0; i < ecount(t); ++i) {
          const int e = efirst(t) + i;
          dz2[e] = arg == i ? dh[t] * elu_grad_from_out(m[e]) : 0.0f;
          db2 += dz2[e];
        }
      }
      outer_acc<EDGES>(gl + GL_W2, dz2, SV, lane);                                 // dW2 += dz2 (x) z1
      atomicAdd(gl + GL_B2 + lane, db2);
      __syncwarp();
#pragma unroll
      for (int e = 0; e < EDGES; ++e) SV[e * H + lane] = dz2[e];
      __syncwarp();
      float dz1[EDGES];
#pragma unroll
      for (int e = 0; e < EDGES; ++e) dz1[e] = 0.0f;
      mv_batch<EDGES>(wl + 5 * H * H, SV, lane, dz1);                              // W2^T dz2
      float dP[NODES], dQ[NODES];
#pragma unroll
      for (int n = 0; n < NODES; ++n) { dP[n] = 0.0f; dQ[n] = 0.0f; }
      float db1 = 0.0f;
#pragma unroll
      for (int e = 0; e < EDGES; ++e) {
        const float pre = P[etgt(e)] + Q[esrc(e)];
        const float d = dz1[e] * (pre > 0.0f ? 1.0f : expf(pre));
        dP[etgt(e)] += d; dQ[esrc(e)] += d;
        db1 += d;
      }
      outer_acc<NODES>(gl + GL_W1, dP, HS, lane);                                  // dW1[:, :32] += dP (x) h
      outer_acc<NODES>(gl + GL_W1 + H * H, dQ, HS, lane);                          // dW1[:, 32:] += dQ (x) h
      atomicAdd(gl + GL_B1 + lane, db1);
      __syncwarp();
#pragma unroll
      for (int n = 0; n < NODES; ++n) { SV[n * H + lane] = dP[n]; SV[(NODES + n) * H + lane] = dQ[n]; }
      __syncwarp();
#pragma unroll
      for (int n = 0; n < NODES; ++n) dh[n] = 0.0f;
      mv_batch<NODES>(wl + 3 * H * H, SV, lane, dh);                               // W1a^T dP
      mv_batch<NODES>(wl + 4 * H * H, SV + NODES * H, lane, dh);                   // + W1b^T dQ
    }
    // ---------------------------------------------------------------- input layers
    {
#pragma unroll
      for (int k = 0; k < 16; ++k) atomicAdd(G + G_IN1 + k * H + lane, dh[0] * OBS[k]);
      atomicAdd(G + GB_IN1 + lane, dh[0]);
      float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f, gb = 0.f;
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const int jj = joint_of(j);
        g0 = fmaf(dh[1 + j], OBS[16 + jj], g0); g1 = fmaf(dh[1 + j], OBS[28 + jj], g1);
        g2 = fmaf(dh[1 + j], OBS[40 + jj], g2); g3 = fmaf(dh[1 + j], OBS[52 + jj], g3);
        gb += dh[1 + j];
      }
      atomicAdd(G + G_IN2 + 0 * H + lane, g0); atomicAdd(G + G_IN2 + 1 * H + lane, g1);
      atomicAdd(G + G_IN2 + 2 * H + lane, g2); atomicAdd(G + G_IN2 + 3 * H + lane, g3);
      atomicAdd(G + GB_IN2 + lane, gb);
    }
  }
  __syncthreads();
  float* out = a.partials + (size_t)blockIdx.x * G_FLOATS;
  for (int i = tid; i < G_FLOATS; i += THREADS) out[i] = G[i];
}

// sum of the per-CTA gradient images in CTA order -> the 20 gradient tensors in nn.Linear layouts
struct ReduceArgs { const float* partials; int num_ctas; float* g[20]; };
__global__ void k_gnn_grad_reduce(const ReduceArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= G_FLOATS) return;
  float s = 0.0f;
  for (int c = 0; c < a.num_ctas; ++c) s += __ldcg(a.partials + (size_t)c * G_FLOATS + j);
  if (j < GB_IN1) { const int k = j / H, c = j % H; a.g[0][c * 16 + k] = s; }
  else if (j < G_IN2) a.g[1][j - GB_IN1] = s;
  else if (j < GB_IN2) { const int r = j - G_IN2, k = r / H, c = r % H; a.g[2][c * 4 + k] = s; }
  else if (j < GL) a.g[3][j - GB_IN2] = s;
  else if (j < G_ACT) {
    const int l = (j - GL) / GL_STRIDE, r = (j - GL) - l * GL_STRIDE;
    if (r < GL_B1) { const int k = r / H, c = r % H; a.g[4 + 4 * l][c * 64 + k] = s; }
    else if (r < GL_W2) a.g[5 + 4 * l][r - GL_B1] = s;
    else if (r < GL_B2) { const int q = r - GL_W2, k = q / H, c = q % H; a.g[6 + 4 * l][c * H + k] = s; }
    else a.g[7 + 4 * l][r - GL_B2] = s;
  }
  else if (j < GB_ACT) a.g[16][j - G_ACT] = s;
  else if (j == GB_ACT) a.g[17][0] = s;
  else if (j < GB_VAL) a.g[18][j - G_VAL] = s;
  else a.g[19][0] = s;
}

// nn.Linear W [32, ld], columns [col0, col0 + K) -> packed forward form [k/4][c][4]
__global__ void k_pack_fwd(const float* __restrict__ w, float* __restrict__ dst, int ld, int col0, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * K) return;
  const int c = i / K, k = i - c * K;
  dst[((k >> 2) * 32 + c) * 4 + (k & 3)] = w[(size_t)c * ld + col0 + k];
}
// the same block -> packed transposed form [c/4][k][4] (K = 32)
__global__ void k_pack_tr(const float* __restrict__ w, float* __restrict__ dst, int ld, int col0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * H) return;
  const int c = i / H, k = i - c * H;
  dst[((c >> 2) * 32 + k) * 4 + (c & 3)] = w[(size_t)c * ld + col0 + k];
}
__global__ void k_copy_n(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

}  // namespace gnn_train
}  // namespace lm

using namespace lm;
using namespace lm::gnn_train;

// same 20 tensors as lm_gnn_pack (called from it)
int lm_gnn_train_pack(LmGnn* g, const float* const* w, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!g->train_blob) {
    if (cudaMalloc(&g->train_blob, BLOB_FLOATS * sizeof(float)) != cudaSuccess) { g->train_blob = nullptr; return lm_fail((int)cudaErrorMemoryAllocation, "lm_gnn_pack: cudaMalloc failed"); }
    cudaMemsetAsync(g->train_blob, 0, BLOB_FLOATS * sizeof(float), s);
  }
  float* B = g->train_blob;
  k_pack_fwd<<<(H * 16 + 127) / 128, 128, 0, s>>>(w[0], B + WP_IN1, 16, 0, 16);
  k_copy_n<<<1, 32, 0, s>>>(w[1], B + B_IN1, H);
  k_pack_fwd<<<(H * 4 + 127) / 128, 128, 0, s>>>(w[2], B + WP_IN2, 4, 0, 4);
  k_copy_n<<<1, 32, 0, s>>>(w[3], B + B_IN2, H);
  for (int l = 0; l < 3; ++l) {
    float* L = B + WL + l * WL_STRIDE;
    const float* w1 = w[4 + 4 * l];
    const float* w2 = w[6 + 4 * l];
    k_pack_fwd<<<(H * H + 127) / 128, 128, 0, s>>>(w1, L + 0 * H * H, 2 * H, 0, H);
    k_pack_fwd<<<(H * H + 127) / 128, 128, 0, s>>>(w1, L + 1 * H * H, 2 * H, H, H);
    k_pack_fwd<<<(H * H + 127) / 128, 128, 0, s>>>(w2, L + 2 * H * H, H, 0, H);
    k_pack_tr<<<(H * H + 127) / 128, 128, 0, s>>>(w1, L + 3 * H * H, 2 * H, 0);
    k_pack_tr<<<(H * H + 127) / 128, 128, 0, s>>>(w1, L + 4 * H * H, 2 * H, H);
    k_pack_tr<<<(H * H + 127) / 128, 128, 0, s>>>(w2, L + 5 * H * H, H, 0);
    k_copy_n<<<1, 32, 0, s>>>(w[5 + 4 * l], B + B_L + l * 2 * H, H);
    k_copy_n<<<1, 32, 0, s>>>(w[7 + 4 * l], B + B_L + l * 2 * H + H, H);
  }
  k_copy_n<<<1, 32, 0, s>>>(w[16], B + W_ACT, H);
  k_copy_n<<<1, 32, 0, s>>>(w[17], B + B_HEAD, 1);
  k_copy_n<<<1, 32, 0, s>>>(w[18], B + W_VAL, H);
  k_copy_n<<<1, 32, 0, s>>>(w[19], B + B_HEAD + 1, 1);
  return lm_check_launch("lm_gnn_pack (training image)");
}

void lm_gnn_train_release(LmGnn* g) {
  if (g->train_blob) cudaFree(g->train_blob);
  g->train_blob = nullptr;
}

extern "C" {

int32_t lm_gnn_grad_workspace_floats(void) { return G_FLOATS; }

int lm_gnn_backward(LmGnn* g, const float* obs, const float* grad_mu, const float* grad_value, int32_t n, float* workspace,
                    int32_t workspace_ctas, float* const* grads, int32_t num_grads, void* stream) {
  if (!g || !obs || !grad_mu || !grad_value || !workspace || !grads || n <= 0 || workspace_ctas <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_gnn_backward: bad argument");
  if (num_grads != 20) return lm_fail(LM_ERR_BAD_ARG, "lm_gnn_backward: expected 20 gradient tensors (the order of lm_gnn_pack)");
  if (!g->ready || !g->train_blob) return lm_fail(LM_ERR_BAD_ARG, "lm_gnn_backward: weights not packed (call lm_gnn_pack)");
  for (int i = 0; i < 20; ++i) if (!grads[i]) return lm_fail(LM_ERR_BAD_ARG, "lm_gnn_backward: gradient tensor %d is NULL", i);
  {
    cudaError_t e = cudaFuncSetAttribute(k_gnn_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return lm_fail((int)e, "lm_gnn_backward: smem attribute: %s", cudaGetErrorString(e));
  }
  const int want = (n + WARPS - 1) / WARPS;
  const int cap = workspace_ctas < g->sm_count ? workspace_ctas : g->sm_count;
  const int grid = want < cap ? want : cap;
  Args a{g->train_blob, obs, grad_mu, grad_value, n, workspace};
  k_gnn_backward<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(a);
  ReduceArgs r;
  r.partials = workspace; r.num_ctas = grid;
  for (int i = 0; i < 20; ++i) r.g[i] = grads[i];
  k_gnn_grad_reduce<<<(G_FLOATS + 255) / 256, 256, 0, (cudaStream_t)stream>>>(r);
  return lm_check_launch("lm_gnn_backward");
}

}  // extern "C"
