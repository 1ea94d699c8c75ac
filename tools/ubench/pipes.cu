// Micro-benchmark: issue cost (cycles per warp instruction per SM sub-partition) of MUFU.EX2, F2FP.BF16.F32.PACK_AB,
// PRMT, and their mixes, with 1..4 warps per sub-partition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#define N_IT 256
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t f2fp(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b)); return r; }
template <int MODE>
__global__ void k(float* out, long long* cyc) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) x[i] = ex2(x[i]);
      if (MODE == 1) { acc ^= f2fp(x[i], x[(i + 1) & 7]); }
      if (MODE == 2) { x[i] = ex2(x[i]); if (i & 1) acc ^= f2fp(x[i], x[i - 1]); }
      if (MODE == 3) { x[i] = ex2(x[i]); acc ^= f2fp(x[i], x[(i + 1) & 7]); }
      if (MODE == 4) { acc ^= prmt(__float_as_uint(x[i]) + 0x8000u, __float_as_uint(x[(i + 1) & 7]) + 0x8000u); }
      if (MODE == 5) { x[i] = ex2(x[i]); if (i & 1) acc ^= prmt(__float_as_uint(x[i]) + 0x8000u, __float_as_uint(x[i - 1]) + 0x8000u); }
      if (MODE == 6) { x[i] = fmaf(x[i], 1.0001f, 0.5f); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const char* names[] = {"MUFU.EX2 x8", "F2FP x8", "EX2 x8 + F2FP x4", "EX2 x8 + F2FP x8", "IADD x16 + PRMT x8", "EX2 x8 + (2 IADD + PRMT) x4", "FFMA x8"};
  for (int mode = 0; mode < 7; ++mode)
    for (int warps = 4; warps <= 16; warps *= 2) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (mode) {
          case 0: k<0><<<1, warps * 32>>>(out, cyc); break; case 1: k<1><<<1, warps * 32>>>(out, cyc); break;
          case 2: k<2><<<1, warps * 32>>>(out, cyc); break; case 3: k<3><<<1, warps * 32>>>(out, cyc); break;
          case 4: k<4><<<1, warps * 32>>>(out, cyc); break; case 5: k<5><<<1, warps * 32>>>(out, cyc); break;
          case 6: k<6><<<1, warps * 32>>>(out, cyc); break;
        }
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-32s warps/SMSP=%d  cycles per 8-element group per warp-slot: %.1f\n", names[mode], warps / 4, (double)h / N_IT / (warps / 4));
    }
  return 0;
}
