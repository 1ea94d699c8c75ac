// Micro-benchmark: what a warp-per-row streaming kernel reaches on HBM, to separate the cost of the LayerNorm
// kernels' structure (row per warp, short CTAs, in-place fp32 stream + bf16 side streams, a warp reduction between
// load and store) from the chip's copy bandwidth.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rowcopy rowcopy.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
constexpr int D = 768, NV = 6;
// MODE 0: flat grid-stride float4 copy; 1: warp per row, out of place; 2: warp per row in place (x = x * 1.0001)
// 3: LN-like: x fp32 in place + bf16 branch in + bf16 out; 4: as 3 plus a warp all-reduce between loads and stores
// 5: as 4 plus gamma / beta / per-sample shift / scale loads after the reduction (cache hits); 6: as 5 plus the runtime
// integer divisions of the row -> sample map in front of the loads; 7: as 6 with the fp32 stream written out of place
__device__ float g_params[4 * 768 + 4096 * 2 * 768];
template <int MODE>
__global__ void __launch_bounds__(256) k(float* __restrict__ x, float* __restrict__ y, const uint2* __restrict__ br, uint2* __restrict__ ob, long long rows, int s0) {
  if (MODE == 0) {
    const long long n4 = rows * D / 4;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL)
      reinterpret_cast<float4*>(y)[i] = reinterpret_cast<const float4*>(x)[i];
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = blockIdx.x * 8LL + warp;
  if (r >= rows) return;
  int sample = 0;
  if (MODE >= 6) { sample = (int)(r / s0) % 4096; if ((int)(r % s0) == 0 && s0 == 12345) return; }
  else if (MODE == 5) sample = (int)(r >> 8) % 4096;
  float4 v[NV]; uint2 b[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = *reinterpret_cast<const float4*>(x + r * D + lane * 4 + 128 * i);
    if (MODE >= 3) b[i] = br[(r * D + lane * 4 + 128 * i) / 4];
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (MODE >= 3) { v[i].x += __uint_as_float(b[i].x << 16); v[i].y += __uint_as_float(b[i].x & 0xffff0000u); v[i].z += __uint_as_float(b[i].y << 16); v[i].w += __uint_as_float(b[i].y & 0xffff0000u); }
    else { v[i].x *= 1.0001f; }
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  float* dst = (MODE == 1 || MODE == 7) ? y : x;
#pragma unroll
  for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(dst + r * D + lane * 4 + 128 * i) = v[i];
  if (MODE == 4) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if (MODE >= 3) {
    const float m = s * (1.f / D);
    if (MODE >= 5) {
      const float* gam = g_params, *bet = g_params + 768, *sh = g_params + 4 * 768 + (long long)sample * 1536, *sc = sh + 768;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane * 4 + 128 * i;
        const float4 a = *reinterpret_cast<const float4*>(gam + c), b2 = *reinterpret_cast<const float4*>(bet + c);
        const float4 c2 = *reinterpret_cast<const float4*>(sh + c), d2 = *reinterpret_cast<const float4*>(sc + c);
        v[i].x = ((v[i].x - m) * a.x + b2.x) * (1.f + d2.x) + c2.x; v[i].y = ((v[i].y - m) * a.y + b2.y) * (1.f + d2.y) + c2.y;
        v[i].z = ((v[i].z - m) * a.z + b2.z) * (1.f + d2.z) + c2.z; v[i].w = ((v[i].w - m) * a.w + b2.w) * (1.f + d2.w) + c2.w;
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[i].x - m, v[i].y - m), p1 = __floats2bfloat162_rn(v[i].z - m, v[i].w - m);
      ob[(r * D + lane * 4 + 128 * i) / 4] = make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
    }
  }
}
#include <cstdlib>
int main(int argc, char** argv) {
  const long long rows = (argc > 1 ? atoll(argv[1]) : 131584LL * 4);   // default 1.6 GB fp32: larger than L2
  float *x, *y; uint2 *br, *ob;
  cudaMalloc(&x, rows * D * 4); cudaMalloc(&y, rows * D * 4); cudaMalloc(&br, rows * D * 2); cudaMalloc(&ob, rows * D * 2);
  cudaMemset(x, 0, rows * D * 4); cudaMemset(br, 0, rows * D * 2);
  const char* names[] = {"flat float4 copy (grid-stride, 148x8 CTAs)", "warp per row, out of place", "warp per row, in place", "LN-like 4 streams (fp32 in place + bf16 in + bf16 out)", "LN-like + warp reduction", "+ gamma/beta/shift/scale loads", "+ row->sample integer divisions", "+ fp32 stream out of place"};
  const double bytes[] = {8.0, 8.0, 8.0, 12.0, 12.0, 12.0, 12.0, 12.0};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 8; ++mode) {
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      const int grid = mode == 0 ? 148 * 8 : (int)((rows + 7) / 8);
      cudaEventRecord(e0);
      switch (mode) {
        case 0: k<0><<<grid, 256>>>(x, y, br, ob, rows, 257); break; case 1: k<1><<<grid, 256>>>(x, y, br, ob, rows, 257); break;
        case 2: k<2><<<grid, 256>>>(x, y, br, ob, rows, 257); break; case 3: k<3><<<grid, 256>>>(x, y, br, ob, rows, 257); break;
        case 4: k<4><<<grid, 256>>>(x, y, br, ob, rows, 257); break; case 5: k<5><<<grid, 256>>>(x, y, br, ob, rows, 257); break;
        case 6: k<6><<<grid, 256>>>(x, y, br, ob, rows, 257); break; case 7: k<7><<<grid, 256>>>(x, y, br, ob, rows, 257); break;
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-58s %7.3f ms  %7.1f GB/s\n", names[mode], best, bytes[mode] * rows * D / best / 1e6);
  }
  return 0;
}
