// Attention dispatch.  The tcgen05/TMEM kernels live in attention_tc.cu; the CUDA-core kernels in
// attention_simt.cu are the in-library checker and cover shapes the tensor-core path does not.
#include "common.cuh"
#include "kernels.cuh"

#include <stdlib.h>
#include <string.h>

namespace umd {

int attention_fwd_tc(const AttnArgs& a, cudaStream_t st);
int attention_bwd_tc(const AttnBwdArgs& a, cudaStream_t st);
bool attention_tc_supported(const RowMap& rm, int nsamples, int H, int Dh);

static int attn_impl() {
  static int impl = -1;
  if (impl < 0) {
    const char* e = getenv("UMD_ATTN_IMPL");  // "simt" forces the CUDA-core checker kernels
    impl = (e && strcmp(e, "simt") == 0) ? 0 : 1;
  }
  return impl;
}

// algorithmic FLOPs: QK^T and PV, 2*S*S*Dh each per (sample, head)
static double attn_flops(const RowMap& rm, int nsamples, int H, int Dh) {
  const double n0 = rm.n0 < nsamples ? rm.n0 : nsamples, n1 = nsamples - n0;
  return 4.0 * Dh * H * (n0 * rm.s0 * static_cast<double>(rm.s0) + n1 * rm.s1 * static_cast<double>(rm.s1));
}

int attention_fwd(const AttnArgs& a, cudaStream_t st) {
  ProfScope prof(PC_ATTN_FWD, attn_flops(a.rm, a.nsamples, a.H, a.Dh), st);
  if (attn_impl() == 1 && attention_tc_supported(a.rm, a.nsamples, a.H, a.Dh)) return attention_fwd_tc(a, st);
  return attention_fwd_simt(a, st);
}
int attention_bwd(const AttnBwdArgs& a, cudaStream_t st) {
  // dV, dP, dQ, dK: 4 contractions + the recomputed QK^T = 2.5x the forward
  ProfScope prof(PC_ATTN_BWD, 2.5 * attn_flops(a.rm, a.nsamples, a.H, a.Dh), st);
  if (attn_impl() == 1 && attention_tc_supported(a.rm, a.nsamples, a.H, a.Dh)) return attention_bwd_tc(a, st);
  return attention_bwd_simt(a, st);
}

}  // namespace umd
