// Host-side error plumbing shared by the translation units of liblm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

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
