// scanner/util/cuda.h -- CU_CHECK / CUDA_PROTECT as op authors use them (reference
// scanner/util/cuda.h:40-48): a failing CUDA call is fatal (the engine treats it as a worker
// failure and re-queues the interval).
#pragma once
#include <cuda_runtime.h>

#include "scanner/util/common.h"

#define CU_CHECK(ans)                                                                     \
  do {                                                                                    \
    cudaError_t code__ = (ans);                                                           \
    if (code__ != cudaSuccess)                                                            \
      LOG(FATAL) << "CUDA error " << (int)code__ << ": " << cudaGetErrorString(code__);   \
  } while (0)

#define CUDA_PROTECT(s) s
