// Shared helpers for the pygho_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pygho_b200.h"

namespace pgh {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

inline int arg_error(const char* what) {
  set_error("bad argument: %s", what);
  return -1;
}

#define PGH_CUDA(expr)                                                   \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) {                                             \
      pgh::set_error("%s: %s", #expr, cudaGetErrorString(_e));           \
      return (int)_e;                                                    \
    }                                                                    \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline unsigned blocks_for(int64_t n, int per_block) {
  return (unsigned)((n + per_block - 1) / per_block);
}

constexpr int kSMs = 148;  // B200

extern int g_tune[8];  // run-time tuning knobs (pgh_set_tuning), defined in seg_gmr.cu

}  // namespace pgh
