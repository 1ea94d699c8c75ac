// Straightforward CUDA-core attention (forward and backward) for one (sample, head) per CTA.
// This is the in-library checker for the tcgen05 attention kernels (attention.cu) and the path
// used for head sizes / sequence lengths those kernels do not cover.  Semantics:
// flax MultiHeadDotProductAttention, vit.py:82-87 — softmax((q/sqrt(Dh)) k^T) v, no mask.
#include "common.cuh"
#include "ptx.cuh"
#include "kernels.cuh"

namespace umd {

extern long long g_launch_count;

constexpr int ATT_DH = 64;
constexpr int ATT_MAX_S = 288;

__device__ __forceinline__ void load_row64(const __nv_bfloat16* p, float (&r)[64], float mul) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 v = *reinterpret_cast<const uint4*>(p + c * 8);
    r[c * 8 + 0] = bf16_lo(v.x) * mul; r[c * 8 + 1] = bf16_hi(v.x) * mul;
    r[c * 8 + 2] = bf16_lo(v.y) * mul; r[c * 8 + 3] = bf16_hi(v.y) * mul;
    r[c * 8 + 4] = bf16_lo(v.z) * mul; r[c * 8 + 5] = bf16_hi(v.z) * mul;
    r[c * 8 + 6] = bf16_lo(v.w) * mul; r[c * 8 + 7] = bf16_hi(v.w) * mul;
  }
}
__device__ __forceinline__ float dot64_smem(const float (&q)[64], const __nv_bfloat16* row) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 v = *reinterpret_cast<const uint4*>(row + c * 8);
    s += q[c * 8 + 0] * bf16_lo(v.x) + q[c * 8 + 1] * bf16_hi(v.x) + q[c * 8 + 2] * bf16_lo(v.y) +
         q[c * 8 + 3] * bf16_hi(v.y) + q[c * 8 + 4] * bf16_lo(v.z) + q[c * 8 + 5] * bf16_hi(v.z) +
         q[c * 8 + 6] * bf16_lo(v.w) + q[c * 8 + 7] * bf16_hi(v.w);
  }
  return s;
}
__device__ __forceinline__ void axpy64_smem(float (&acc)[64], float a, const __nv_bfloat16* row) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 v = *reinterpret_cast<const uint4*>(row + c * 8);
    acc[c * 8 + 0] += a * bf16_lo(v.x); acc[c * 8 + 1] += a * bf16_hi(v.x);
    acc[c * 8 + 2] += a * bf16_lo(v.y); acc[c * 8 + 3] += a * bf16_hi(v.y);
    acc[c * 8 + 4] += a * bf16_lo(v.z); acc[c * 8 + 5] += a * bf16_hi(v.z);
    acc[c * 8 + 6] += a * bf16_lo(v.w); acc[c * 8 + 7] += a * bf16_hi(v.w);
  }
}
__device__ __forceinline__ void store_row64(__nv_bfloat16* p, const float (&r)[64]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    *reinterpret_cast<uint4*>(p + c * 8) =
        make_uint4(pack_bf16x2(r[c * 8], r[c * 8 + 1]), pack_bf16x2(r[c * 8 + 2], r[c * 8 + 3]),
                   pack_bf16x2(r[c * 8 + 4], r[c * 8 + 5]), pack_bf16x2(r[c * 8 + 6], r[c * 8 + 7]));
  }
}

__global__ void __launch_bounds__(128) attn_fwd_simt_kernel(AttnArgs a) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int h = blockIdx.x, n = blockIdx.y;
  const int S = seq_of(a.rm, n);
  const long long row0 = row_of(a.rm, n, 0);
  const int ld = 3 * a.H * ATT_DH;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smraw);
  __nv_bfloat16* sV = sK + S * ATT_DH;
  for (int i = threadIdx.x; i < S * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const __nv_bfloat16* src = a.qkv + (row0 + r) * ld + h * ATT_DH + c * 8;
    *reinterpret_cast<uint4*>(sK + r * ATT_DH + c * 8) = *reinterpret_cast<const uint4*>(src + a.H * ATT_DH);
    *reinterpret_cast<uint4*>(sV + r * ATT_DH + c * 8) = *reinterpret_cast<const uint4*>(src + 2 * a.H * ATT_DH);
  }
  __syncthreads();
  const float sl2 = a.scale * 1.4426950408889634f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    float q[64];
    load_row64(a.qkv + (row0 + i) * ld + h * ATT_DH, q, sl2);
    float m = -INFINITY;
    for (int j = 0; j < S; ++j) m = fmaxf(m, dot64_smem(q, sK + j * ATT_DH));
    float l = 0.f;
    float acc[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) acc[d] = 0.f;
    for (int j = 0; j < S; ++j) {
      const float p = exp2f(dot64_smem(q, sK + j * ATT_DH) - m);
      l += p;
      axpy64_smem(acc, p, sV + j * ATT_DH);
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < 64; ++d) acc[d] *= inv;
    store_row64(a.out + (row0 + i) * (a.H * ATT_DH) + h * ATT_DH, acc);
    if (a.lse) a.lse[(row0 + i) * a.H + h] = (m + log2f(l)) * 0.6931471805599453f;
  }
}

// Backward: pass A thread-per-query -> dQ; pass B thread-per-key -> dK, dV.
__global__ void __launch_bounds__(128) attn_bwd_simt_kernel(AttnBwdArgs a) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int h = blockIdx.x, n = blockIdx.y;
  const int S = seq_of(a.rm, n);
  const long long row0 = row_of(a.rm, n, 0);
  const int D = a.H * ATT_DH;
  const int ld = 3 * D;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smraw);
  __nv_bfloat16* sK = sQ + S * ATT_DH;
  __nv_bfloat16* sV = sK + S * ATT_DH;
  __nv_bfloat16* sdO = sV + S * ATT_DH;
  float* sLse = reinterpret_cast<float*>(sdO + S * ATT_DH);
  float* sDelta = sLse + S;
  for (int i = threadIdx.x; i < S * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const __nv_bfloat16* src = a.qkv + (row0 + r) * ld + h * ATT_DH + c * 8;
    *reinterpret_cast<uint4*>(sQ + r * ATT_DH + c * 8) = *reinterpret_cast<const uint4*>(src);
    *reinterpret_cast<uint4*>(sK + r * ATT_DH + c * 8) = *reinterpret_cast<const uint4*>(src + D);
    *reinterpret_cast<uint4*>(sV + r * ATT_DH + c * 8) = *reinterpret_cast<const uint4*>(src + 2 * D);
    *reinterpret_cast<uint4*>(sdO + r * ATT_DH + c * 8) =
        *reinterpret_cast<const uint4*>(a.dout + (row0 + r) * D + h * ATT_DH + c * 8);
  }
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    float o[64], g[64];
    load_row64(a.out + (row0 + i) * D + h * ATT_DH, o, 1.f);
    load_row64(a.dout + (row0 + i) * D + h * ATT_DH, g, 1.f);
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < 64; ++k) d += o[k] * g[k];
    sDelta[i] = d;
    sLse[i] = a.lse[(row0 + i) * a.H + h];
  }
  __syncthreads();
  // pass A: dQ_i = scale * sum_j p_ij (dp_ij - delta_i) k_j
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    float q[64], g[64], dq[64];
    load_row64(sQ + i * ATT_DH, q, a.scale);
    load_row64(sdO + i * ATT_DH, g, 1.f);
#pragma unroll
    for (int d = 0; d < 64; ++d) dq[d] = 0.f;
    const float lse = sLse[i], delta = sDelta[i];
    for (int j = 0; j < S; ++j) {
      const float p = __expf(dot64_smem(q, sK + j * ATT_DH) - lse);
      const float dp = dot64_smem(g, sV + j * ATT_DH);
      axpy64_smem(dq, p * (dp - delta) * a.scale, sK + j * ATT_DH);
    }
    store_row64(a.dqkv + (row0 + i) * ld + h * ATT_DH, dq);
  }
  // pass B: dV_j = sum_i p_ij dO_i ; dK_j = scale * sum_i p_ij (dp_ij - delta_i) q_i
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    float kj[64], vj[64];
    load_row64(sK + j * ATT_DH, kj, a.scale);
    load_row64(sV + j * ATT_DH, vj, 1.f);
    float dk[64], dv[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) dk[d] = dv[d] = 0.f;
    for (int i = 0; i < S; ++i) {
      const float p = __expf(dot64_smem(kj, sQ + i * ATT_DH) - sLse[i]);
      const float dp = dot64_smem(vj, sdO + i * ATT_DH);
      axpy64_smem(dv, p, sdO + i * ATT_DH);
      axpy64_smem(dk, p * (dp - sDelta[i]) * a.scale, sQ + i * ATT_DH);
    }
    store_row64(a.dqkv + (row0 + j) * ld + D + h * ATT_DH, dk);
    store_row64(a.dqkv + (row0 + j) * ld + 2 * D + h * ATT_DH, dv);
  }
}

static int max_seq(const RowMap& rm, int nsamples) {
  int s = rm.n0 > 0 ? rm.s0 : 0;
  if (nsamples > rm.n0 && rm.s1 > s) s = rm.s1;
  return s;
}

int attention_fwd_simt(const AttnArgs& a, cudaStream_t st) {
  if (a.nsamples <= 0) return UMD_OK;
  const int S = max_seq(a.rm, a.nsamples);
  UMD_REQUIRE(a.Dh == ATT_DH && S <= ATT_MAX_S, "attention: head dim %d / sequence %d unsupported", a.Dh, S);
  const int smem = 2 * S * ATT_DH * 2;
  static bool cfg = false;
  if (!cfg) {
    UMD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * ATT_MAX_S * ATT_DH * 2));
    cfg = true;
  }
  attn_fwd_simt_kernel<<<dim3(a.H, a.nsamples), 128, smem, st>>>(a);
  ++g_launch_count;
  UMD_CHECK_CUDA(cudaGetLastError());
  return UMD_OK;
}

int attention_bwd_simt(const AttnBwdArgs& a, cudaStream_t st) {
  if (a.nsamples <= 0) return UMD_OK;
  const int S = max_seq(a.rm, a.nsamples);
  UMD_REQUIRE(a.Dh == ATT_DH && S <= ATT_MAX_S, "attention: head dim %d / sequence %d unsupported", a.Dh, S);
  const int smem = 4 * S * ATT_DH * 2 + 2 * S * 4;
  static bool cfg = false;
  if (!cfg) {
    UMD_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        4 * ATT_MAX_S * ATT_DH * 2 + 2 * ATT_MAX_S * 4));
    cfg = true;
  }
  attn_bwd_simt_kernel<<<dim3(a.H, a.nsamples), 128, smem, st>>>(a);
  ++g_launch_count;
  UMD_CHECK_CUDA(cudaGetLastError());
  return UMD_OK;
}

}  // namespace umd
