// hist_csa.cuh -- bit-sliced streaming histogram (placeholder until the kernel lands).
#pragma once
#include "scn_common.cuh"
namespace scn { namespace csa {
inline bool eligible(const uint8_t* const*, int, size_t) { return false; }
inline int launch(const uint8_t* const*, int, size_t, int32_t*, cudaStream_t) {
  return SCN_E_UNSUPPORTED;
}
}}  // namespace scn::csa
