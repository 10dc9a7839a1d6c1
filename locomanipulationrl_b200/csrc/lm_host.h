// Host-side error plumbing shared by the translation units of liblm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <utility>

inline char* lm_error_buffer() {
  static thread_local char buf[512] = "";
  return buf;
}
inline void lm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(lm_error_buffer(), 512, fmt, ap);
  va_end(ap);
}
inline int lm_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(lm_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}
// launch-configuration errors only (no synchronisation)
inline int lm_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  return lm_fail((int)e, "%s: %s", what, cudaGetErrorString(e));
}

// Launch with (pdl = true) or without the programmatic-stream-serialization attribute; LM_NO_PDL=1 switches it off globally.
inline bool lm_pdl_enabled() {
  static const bool on = !(getenv("LM_NO_PDL") && getenv("LM_NO_PDL")[0] == '1');
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t lm_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && lm_pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
