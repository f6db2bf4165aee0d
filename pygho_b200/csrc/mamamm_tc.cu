// tcgen05 / TMA tensor-core path of the 2-FWL contraction (placeholder until the UMMA
// kernel lands; algo=1 reports "not available" so that callers fall back explicitly).
#include "common.cuh"

namespace pgh {
int mamamm_tc_launch(const float*, int, const float*, int, const unsigned char*, int64_t,
                     int64_t, int64_t, int64_t, int64_t, float*, cudaStream_t) {
  set_error("mamamm algo=1 (tcgen05) is not available in this build");
  return -2;
}
}  // namespace pgh
