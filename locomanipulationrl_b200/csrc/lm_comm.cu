// Host side of the one-shot NVLink all-reduce (lm_comm.cuh): receive memory, cudaIpc handle exchange, the stand-alone
// all-reduce launch.  One process per GPU; the handles travel between the processes through the caller's own channel
// (torch.distributed.all_gather of 64-byte blobs in locomanipulationrl_b200/dist.py).
#include <string.h>

#include "lm_comm.cuh"
#include "lm_host.h"

using namespace lm;

struct LmComm {
  int device, rank, world, sm_count;
  long long slot_floats;
  void* region;                       // flags + data, cudaMalloc'ed here
  void* peer_region[comm::MAX_WORLD]; // mapped peers (NULL for self / not opened)
  unsigned int* ctl;
  comm::Ctx ctx;
  bool opened;
};

static constexpr size_t FLAG_BYTES = 256;      // MAX_WORLD flags, padded so that the data region starts 256-byte aligned

static void fill_ctx(LmComm* c) {
  comm::Ctx& x = c->ctx;
  x.rank = c->rank; x.world = c->world; x.slot_floats = c->slot_floats;
  for (int p = 0; p < comm::MAX_WORLD; ++p) {
    void* base = p == c->rank ? c->region : c->peer_region[p];
    x.peer_flags[p] = reinterpret_cast<unsigned int*>(base);
    x.peer_data[p] = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + FLAG_BYTES) : nullptr;
  }
  x.local_flags = reinterpret_cast<unsigned int*>(c->region);
  x.local_data = reinterpret_cast<float*>(reinterpret_cast<char*>(c->region) + FLAG_BYTES);
  x.ctl = c->ctl;
}

const comm::Ctx* lm_comm_ctx(const LmComm* c) { return &c->ctx; }
long long lm_comm_slot_floats(const LmComm* c) { return c->slot_floats; }
int lm_comm_max_ctas(const LmComm* c) { return c->sm_count; }

namespace lm {
namespace comm {

// stand-alone all-reduce (sum) of n floats, in place.  grid <= SM count (every CTA must be resident: they wait for each other's ranks)
// Optional gather: elements [g.off, g.off + g.n_f32 + g.n_i64) of the message are taken from two other places (fp32 values, then
// int64 counters converted to fp32) instead of `data` — the episode statistics ride in the gradient arena's tail without a
// separate packing launch.
struct Gather { long long off; const float* f32; int n_f32; const long long* i64; int n_i64; float f32_scale; float one_more; int has_one_more; };
__global__ void __launch_bounds__(1024, 1) k_allreduce(const Ctx c, float* __restrict__ data, long long n, const Gather g) {
  const unsigned int epoch = epoch_of(c);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < n; f += stride) {
    const long long r = f - g.off;
    float v;
    if (r >= 0 && r < g.n_f32) v = g.f32[r] * g.f32_scale;
    else if (r >= g.n_f32 && r < g.n_f32 + g.n_i64) v = (float)g.i64[r - g.n_f32];
    else if (g.has_one_more && r == g.n_f32 + g.n_i64) v = g.one_more;
    else v = data[f];
    push(c, epoch, f, v);
  }
  if (signal_and_wait(c, epoch)) {
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < n; f += stride) data[f] = sum(c, epoch, f);
  }
  leave(c);
}

}  // namespace comm
}  // namespace lm

extern "C" {

int lm_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

LmComm* lm_comm_create(int device, int rank, int world, int64_t max_floats) {
  if (world < 1 || world > comm::MAX_WORLD || rank < 0 || rank >= world || max_floats <= 0) {
    lm_set_error("lm_comm_create: need 1 <= world <= %d, 0 <= rank < world, max_floats > 0", comm::MAX_WORLD);
    return nullptr;
  }
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { lm_set_error("lm_comm_create: %s", cudaGetErrorString(e)); return nullptr; }
  cudaSetDevice(device);
  LmComm* c = new LmComm();
  memset(c, 0, sizeof(*c));
  c->device = device; c->rank = rank; c->world = world; c->sm_count = prop.multiProcessorCount;
  c->slot_floats = (max_floats + 63) & ~63ll;
  const size_t bytes = FLAG_BYTES + (size_t)2 * world * c->slot_floats * sizeof(float);
  if (cudaMalloc(&c->region, bytes) != cudaSuccess || cudaMalloc(&c->ctl, 4 * sizeof(unsigned int)) != cudaSuccess) {
    lm_set_error("lm_comm_create: cudaMalloc failed");
    cudaFree(c->region); delete c; return nullptr;
  }
  cudaMemset(c->region, 0, bytes);
  cudaMemset(c->ctl, 0, 4 * sizeof(unsigned int));
  cudaDeviceSynchronize();
  c->opened = world == 1;
  fill_ctx(c);
  return c;
}

int lm_comm_get_handle(LmComm* c, void* handle_out) {
  if (!c || !handle_out) return lm_fail(LM_ERR_BAD_ARG, "lm_comm_get_handle: NULL argument");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, c->region);
  if (e != cudaSuccess) return lm_fail((int)e, "lm_comm_get_handle: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  memcpy(handle_out, &h, sizeof(h));
  return LM_OK;
}

// handles: world * lm_comm_handle_bytes() bytes in rank order (this rank's own entry is ignored)
int lm_comm_open(LmComm* c, const void* handles) {
  if (!c || !handles) return lm_fail(LM_ERR_BAD_ARG, "lm_comm_open: NULL argument");
  cudaSetDevice(c->device);
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank || c->peer_region[p]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, reinterpret_cast<const char*>(handles) + (size_t)p * sizeof(h), sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(&c->peer_region[p], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      c->peer_region[p] = nullptr;
      return lm_fail((int)e, "lm_comm_open: cudaIpcOpenMemHandle(rank %d): %s (peer access / IPC between the ranks' GPUs is required)", p,
                     cudaGetErrorString(e));
    }
  }
  c->opened = true;
  fill_ctx(c);
  return LM_OK;
}

void lm_comm_destroy(LmComm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  for (int p = 0; p < comm::MAX_WORLD; ++p)
    if (c->peer_region[p]) cudaIpcCloseMemHandle(c->peer_region[p]);
  cudaFree(c->region); cudaFree(c->ctl);
  delete c;
}

static int allreduce_impl(LmComm* c, float* data, int64_t n, const comm::Gather& g, void* stream) {
  if (!c || !data || n <= 0) return lm_fail(LM_ERR_BAD_ARG, "lm_comm_allreduce: bad argument");
  if (!c->opened) return lm_fail(LM_ERR_BAD_ARG, "lm_comm_allreduce: peers not opened (lm_comm_open)");
  if (n > c->slot_floats) return lm_fail(LM_ERR_BAD_ARG, "lm_comm_allreduce: %lld floats exceed the slot size %lld", (long long)n, c->slot_floats);
  const long long want = (n + 1023) / 1024;
  const int grid = (int)(want < c->sm_count ? want : c->sm_count);
  comm::k_allreduce<<<grid, 1024, 0, (cudaStream_t)stream>>>(c->ctx, data, (long long)n, g);
  return lm_check_launch("lm_comm_allreduce");
}

int lm_comm_allreduce(LmComm* c, float* data, int64_t n, void* stream) {
  const comm::Gather g{0, nullptr, 0, nullptr, 0, 1.0f, 0.0f, 0};
  return allreduce_impl(c, data, n, g, stream);
}

int lm_comm_allreduce_gather(LmComm* c, float* data, int64_t n, int64_t gather_off, const float* f32_src, int32_t n_f32, float f32_scale,
                             const int64_t* i64_src, int32_t n_i64, float one_more, void* stream) {
  if (gather_off < 0 || n_f32 < 0 || n_i64 < 0 || gather_off + n_f32 + n_i64 + 1 > n || (n_f32 > 0 && !f32_src) || (n_i64 > 0 && !i64_src))
    return lm_fail(LM_ERR_BAD_ARG, "lm_comm_allreduce_gather: bad gather range");
  const comm::Gather g{(long long)gather_off, f32_src, n_f32, reinterpret_cast<const long long*>(i64_src), n_i64, f32_scale, one_more, 1};
  return allreduce_impl(c, data, n, g, stream);
}

// error word of the handle: 0 = fine, 1 + r = timed out waiting for rank r (host synchronisation on the device)
int lm_comm_status(LmComm* c) {
  if (!c) return -1;
  unsigned int w[4];
  if (cudaMemcpy(w, c->ctl, sizeof(w), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)w[3];
}

}  // extern "C"
