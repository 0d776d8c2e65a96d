// Shared host-side helpers for libhd_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstring>
#include "../../include/hd_b200.h"

namespace hd {

extern std::atomic<long long> g_launches;
void set_last_error(const char *what, cudaError_t e);
void set_last_error_text(const char *what);

inline int check_launch(const char *what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_last_error(what, e);
    return HD_ERR_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return HD_OK;
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace hd

#define HD_REQUIRE(cond, msg)                \
  do {                                       \
    if (!(cond)) {                           \
      hd::set_last_error_text(msg);          \
      return HD_ERR_INVALID;                 \
    }                                        \
  } while (0)
