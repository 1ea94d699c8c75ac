// tcgen05/TMEM fused attention (forward, backward).  Placeholder until the kernels land: reports
// "unsupported" so that the dispatcher in attention.cu uses the CUDA-core kernels.
#include "common.cuh"
#include "kernels.cuh"

namespace umd {
bool attention_tc_supported(const RowMap&, int, int, int) { return false; }
int attention_fwd_tc(const AttnArgs&, cudaStream_t) {
  set_error("attention_fwd_tc: not built");
  return UMD_ERR_UNSUPPORTED;
}
int attention_bwd_tc(const AttnBwdArgs&, cudaStream_t) {
  set_error("attention_bwd_tc: not built");
  return UMD_ERR_UNSUPPORTED;
}
}  // namespace umd
