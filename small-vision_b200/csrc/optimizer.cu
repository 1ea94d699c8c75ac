// Optimiser over the flat parameter arena (train_ae.py:148-151,365-374):
//   clip_by_global_norm(c) -> scale_by_adam(b1, b2, eps=1e-8, mu bf16) -> add_decayed_weights(wd, mask)
//   -> scale by -lr -> apply_updates, plus l2_params / l2_updates and the optional EMA.
// One sum-of-squares reduction and one fused multi-tensor pass; HBM-bound (~26 B / parameter).
#include "common.cuh"
#include "ptx.cuh"
#include "kernels.cuh"

namespace umd {

extern long long g_launch_count;

__device__ __forceinline__ float block_sum_256(float v, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 8) {
    t = sm[threadIdx.x];
    t += __shfl_xor_sync(0xffu, t, 4);
    t += __shfl_xor_sync(0xffu, t, 2);
    t += __shfl_xor_sync(0xffu, t, 1);
  }
  __syncthreads();
  return t;  // valid in thread 0
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n4, float* __restrict__ partials) {
  __shared__ float sm[8];
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  float t = block_sum_256(acc, sm);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int n, float* __restrict__ out,
                                                              int take_sqrt) {
  __shared__ double sm[256];
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) t += static_cast<double>(partials[i]);
  sm[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = static_cast<float>(take_sqrt ? sqrt(sm[0]) : sm[0]);
}

int sumsq(const float* x, long long n, float* partials, int max_partials, float* out, cudaStream_t st) {
  UMD_REQUIRE(n % 4 == 0, "sumsq: count must be a multiple of 4");
  int grid = static_cast<int>(ceil_div_ll(n / 4, 256 * 8));
  if (grid > max_partials) grid = max_partials;
  if (grid < 1) grid = 1;
  sumsq_kernel<<<grid, 256, 0, st>>>(x, n / 4, partials);
  ++g_launch_count;
  UMD_CHECK_CUDA(cudaGetLastError());
  reduce_partials_kernel<<<1, 256, 0, st>>>(partials, grid, out, 0);
  ++g_launch_count;
  UMD_CHECK_CUDA(cudaGetLastError());
  return UMD_OK;
}

struct AdamwParams {
  float* p;
  const float* g;
  __nv_bfloat16* mu;
  float* nu;
  __nv_bfloat16* p_bf16;
  float* ema;
  const uint8_t* wd_flags;   // one flag per 64 consecutive elements
  const float* gnorm_sq;     // device scalar
  float* part_u;
  float* part_p;
  long long n4;
  float clip, lr, b1, b2, eps, wd, bc1, bc2, ema_decay;
};

__global__ void __launch_bounds__(256) adamw_kernel(AdamwParams a) {
  __shared__ float sm[8];
  const float gnorm = sqrtf(*a.gnorm_sq);
  const bool clip = !(gnorm < a.clip);
  float su = 0.f, sp = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < a.n4; i += static_cast<long long>(gridDim.x) * 256) {
    float4 g4 = reinterpret_cast<const float4*>(a.g)[i];
    float4 p4 = reinterpret_cast<float4*>(a.p)[i];
    float4 nu4 = reinterpret_cast<float4*>(a.nu)[i];
    uint2 mu2 = reinterpret_cast<uint2*>(a.mu)[i];
    const bool decay = a.wd_flags[i >> 4] != 0;
    float g[4] = {g4.x, g4.y, g4.z, g4.w};
    float p[4] = {p4.x, p4.y, p4.z, p4.w};
    float nu[4] = {nu4.x, nu4.y, nu4.z, nu4.w};
    float mu[4] = {bf16_lo(mu2.x), bf16_hi(mu2.x), bf16_lo(mu2.y), bf16_hi(mu2.y)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gc = clip ? (g[k] / gnorm) * a.clip : g[k];
      mu[k] = (1.f - a.b1) * gc + a.b1 * mu[k];
      nu[k] = (1.f - a.b2) * (gc * gc) + a.b2 * nu[k];
      float u = (mu[k] / a.bc1) / (sqrtf(nu[k] / a.bc2) + a.eps);
      if (decay) u += a.wd * p[k];
      u = -a.lr * u;
      p[k] += u;
      su += u * u;
      sp += p[k] * p[k];
    }
    reinterpret_cast<float4*>(a.p)[i] = make_float4(p[0], p[1], p[2], p[3]);
    reinterpret_cast<float4*>(a.nu)[i] = make_float4(nu[0], nu[1], nu[2], nu[3]);
    reinterpret_cast<uint2*>(a.mu)[i] = make_uint2(pack_bf16x2(mu[0], mu[1]), pack_bf16x2(mu[2], mu[3]));
    if (a.p_bf16) reinterpret_cast<uint2*>(a.p_bf16)[i] = make_uint2(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]));
    if (a.ema) {
      float4 e = reinterpret_cast<float4*>(a.ema)[i];
      const float d = a.ema_decay;
      e.x = d * p[0] + (1.f - d) * e.x; e.y = d * p[1] + (1.f - d) * e.y;
      e.z = d * p[2] + (1.f - d) * e.z; e.w = d * p[3] + (1.f - d) * e.w;
      reinterpret_cast<float4*>(a.ema)[i] = e;
    }
  }
  float tu = block_sum_256(su, sm);
  float tp = block_sum_256(sp, sm);
  if (threadIdx.x == 0) {
    a.part_u[blockIdx.x] = tu;
    a.part_p[blockIdx.x] = tp;
  }
}

}  // namespace umd

using namespace umd;

extern "C" int umd_sumsq(const float* x, long long n, float* scratch, int scratch_floats, float* out, umd_stream_t stream) {
  return sumsq(x, n, scratch, scratch_floats, out, static_cast<cudaStream_t>(stream));
}

extern "C" int umd_adamw_step(const umd_adamw_args* a, umd_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UMD_REQUIRE(a && a->n > 0 && a->n % 64 == 0, "umd_adamw_step: arena size must be a positive multiple of 64");
  UMD_REQUIRE(a->scratch_floats >= 3 * 1024 + 8, "umd_adamw_step: scratch too small (need >= 3080 floats)");
  // algorithmic bytes per parameter: g r 4 (twice: norm + update), p rw 8, mu rw 4, nu rw 8, bf16 shadow w 2, ema rw 8
  ProfScope prof(PC_OPTIMIZER, static_cast<double>(a->n) * (8 + 8 + 4 + 8 + (a->params_bf16 ? 2 : 0) + (a->ema ? 8 : 0)), st);
  float* part_g = a->scratch;
  float* part_u = a->scratch + 1024;
  float* part_p = a->scratch + 2048;
  float* gnorm_sq = a->scratch + 3072;
  // measurements: [0] l2_params  [1] l2_updates  [2] grad_norm
  UMD_TRY(sumsq(a->grads, a->n, part_g, 1024, gnorm_sq, st));
  AdamwParams k;
  k.p = a->params; k.g = a->grads; k.mu = reinterpret_cast<__nv_bfloat16*>(a->mu); k.nu = a->nu;
  k.p_bf16 = reinterpret_cast<__nv_bfloat16*>(a->params_bf16); k.ema = a->ema; k.wd_flags = a->wd_flags;
  k.gnorm_sq = gnorm_sq; k.part_u = part_u; k.part_p = part_p; k.n4 = a->n / 4;
  k.clip = a->clip_norm; k.lr = a->lr; k.b1 = a->b1; k.b2 = a->b2; k.eps = a->eps; k.wd = a->wd;
  k.bc1 = a->bias_corr1; k.bc2 = a->bias_corr2; k.ema_decay = a->ema_decay;
  int grid = static_cast<int>(ceil_div_ll(k.n4, 256 * 4));
  if (grid > 1024) grid = 1024;
  adamw_kernel<<<grid, 256, 0, st>>>(k);
  ++g_launch_count;
  UMD_CHECK_CUDA(cudaGetLastError());
  reduce_partials_kernel<<<1, 256, 0, st>>>(part_p, grid, a->measurements + 0, 1);
  reduce_partials_kernel<<<1, 256, 0, st>>>(part_u, grid, a->measurements + 1, 1);
  reduce_partials_kernel<<<1, 256, 0, st>>>(gnorm_sq, 1, a->measurements + 2, 1);
  g_launch_count += 3;
  UMD_CHECK_CUDA(cudaGetLastError());
  return UMD_OK;
}
