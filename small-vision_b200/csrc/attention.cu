// Attention dispatch.  The product entry points run the tcgen05 / TMEM kernels of attention_tc.cu and nothing else: a
// shape they do not cover (head dim != 64, more than 272 tokens per sample) is an error, not a slower path.  The CUDA-core
// kernels of attention_simt.cu are an in-library checker, reachable only through the explicit umd_attention_*_simt exports.
#include "common.cuh"
#include "kernels.cuh"

namespace umd {

int attention_fwd_tc(const AttnArgs& a, cudaStream_t st);
int attention_bwd_tc(const AttnBwdArgs& a, cudaStream_t st);
bool attention_tc_supported(const RowMap& rm, int nsamples, int H, int Dh);

// algorithmic FLOPs: QK^T and PV, 2*S*S*Dh each per (sample, head)
static double attn_flops(const RowMap& rm, int nsamples, int H, int Dh) {
  const double n0 = rm.n0 < nsamples ? rm.n0 : nsamples, n1 = nsamples - n0;
  return 4.0 * Dh * H * (n0 * rm.s0 * static_cast<double>(rm.s0) + n1 * rm.s1 * static_cast<double>(rm.s1));
}

static int unsupported(const RowMap& rm, int nsamples, int H, int Dh) {
  set_error("attention: unsupported shape (heads %d of dim %d, %d samples, %d / %d tokens per sample): the tcgen05 kernels "
            "need head dim 64 and 1..272 tokens per sample", H, Dh, nsamples, rm.s0, rm.s1);
  return UMD_ERR_UNSUPPORTED;
}

int attention_fwd(const AttnArgs& a, cudaStream_t st) {
  if (!attention_tc_supported(a.rm, a.nsamples, a.H, a.Dh)) return unsupported(a.rm, a.nsamples, a.H, a.Dh);
  ProfScope prof(PC_ATTN_FWD, attn_flops(a.rm, a.nsamples, a.H, a.Dh), st);
  return attention_fwd_tc(a, st);
}
int attention_bwd(const AttnBwdArgs& a, cudaStream_t st) {
  if (!attention_tc_supported(a.rm, a.nsamples, a.H, a.Dh)) return unsupported(a.rm, a.nsamples, a.H, a.Dh);
  // dV, dP, dQ, dK: 4 contractions + the recomputed QK^T = 2.5x the forward
  ProfScope prof(PC_ATTN_BWD, 2.5 * attn_flops(a.rm, a.nsamples, a.H, a.Dh), st);
  return attention_bwd_tc(a, st);
}

}  // namespace umd
