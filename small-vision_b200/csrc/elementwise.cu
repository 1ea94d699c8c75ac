// Bandwidth-bound kernels of the UMD step: LayerNorm(+adaLN modulate) forward/backward,
// gate backward, column sums, conditioning path, patch embedding (gather-then-embed),
// decoder-input build, q_sample, stable mask argsort, un-patchify and the fused loss.
// All are coalesced, 128-bit vectorised, warp-shuffle / shared-memory reductions; none of
// them reshapes work into a GEMM.  Reference lines are cited per kernel.
#include "common.cuh"
#include "ptx.cuh"
#include "kernels.cuh"

#include <stdlib.h>

namespace umd {

extern long long g_launch_count;

#define UMD_LAUNCH_CHECK()                 \
  do {                                     \
    ++g_launch_count;                      \
    UMD_CHECK_CUDA(cudaGetLastError());    \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float silu_grad_f(float x) {
  float s = 1.f / (1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}

__device__ __forceinline__ void store_row4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void store_row4(__nv_bfloat16* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}
__device__ __forceinline__ float4 load_row4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load_row4(const __nv_bfloat16* p) {
  uint2 v = *reinterpret_cast<const uint2*>(p);
  return make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
}

// =========================================================================================
// LayerNorm (+ adaLN modulate) forward.  nn.LayerNorm eps=1e-6, fast variance
// (vit.py:78-80,96-98,163; modulate vit.py:13-16; final modulation ae.py:166-170).
// One warp per output row; lane owns float4 columns lane*4 + 128*i.
// =========================================================================================
// Work decomposition: grid (sample, row chunk), one warp per row, a warp walks over its rows of the chunk.  All rows of
// a CTA belong to one sample, so the per-column coefficients are combined once per CTA into shared memory:
//   y = xhat * A + B,   A = gamma (1 + scale[n]),   B = beta (1 + scale[n]) + shift[n],   G = gate[n]
// A row then costs 12 + 6 shared-memory reads per thread instead of 24 + 6 cached global loads (gamma, beta, shift,
// scale, gate), which had the L1 data path at 65 % (ncu) and capped the kernel at ~4.4 TB/s (tools/ubench/rowcopy.cu:
// the same row-per-warp streams without the coefficient loads reach 6.8 TB/s), and no row -> sample division is left.
#ifndef LN_FWD_MIN_CTAS
#define LN_FWD_MIN_CTAS 4
#endif
// XinT / XoutT: dtype of the residual stream read / written (float, or __nv_bfloat16 with umd_model_cfg.residual_bf16: the
// reference's own dtype_mm="bfloat16" flow, ae.py:51,100; statistics stay fp32 either way).
template <int NV, typename OutT, typename XinT, typename XoutT>
__global__ void __launch_bounds__(256, LN_FWD_MIN_CTAS) ln_mod_fwd_kernel(LnFwdArgs a, int rows_per_chunk) {
  constexpr int D = NV * 128;
  __shared__ __align__(16) float sA[D];
  __shared__ __align__(16) float sB[D];
  __shared__ __align__(16) float sG[D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x;
  const int S = a.gather_L > 0 ? a.gather_L : seq_of(a.rm, n);
  const int tok_begin = blockIdx.y * rows_per_chunk;
  if (tok_begin >= S) return;
  const int tok_end = min(S, tok_begin + rows_per_chunk);
  {
    const float* shp = a.shift ? a.shift + static_cast<long long>(n) * a.ldmod : nullptr;
    const float* scp = a.scale ? a.scale + static_cast<long long>(n) * a.ldmod : nullptr;
    const float* gp = (a.res_branch && a.res_gate) ? a.res_gate + static_cast<long long>(n) * a.ldgate : nullptr;
    for (int c = threadIdx.x; c < D; c += 256) {
      const float sc1 = scp ? 1.f + scp[c] : 1.f;
      sA[c] = a.gamma[c] * sc1;
      sB[c] = a.beta[c] * sc1 + (shp ? shp[c] : 0.f);
      sG[c] = gp ? gp[c] : 1.f;
    }
  }
  __syncthreads();
  for (int tok = tok_begin + warp; tok < tok_end; tok += 8) {
    int r, in_row;
    if (a.gather_L > 0) {
      r = n * a.gather_L + tok;
      in_row = row_of(a.rm, n, a.gather_off + tok);
    } else {
      r = in_row = row_of(a.rm, n, tok);
    }
    const bool is_cond = a.cond_row && (a.gather_L > 0 ? a.gather_off + tok : tok) == 0;
    const XinT* src = reinterpret_cast<const XinT*>(a.x) + static_cast<long long>(in_row) * D;
    const float* csrc = is_cond ? a.cond_row + static_cast<long long>(n) * D : nullptr;
    const __nv_bfloat16* bp = (a.res_branch && !is_cond) ? a.res_branch + static_cast<long long>(in_row) * D : nullptr;
    // every HBM load of the row is issued before the first store: x_out may alias x, so a store in between would
    // pin all later loads behind it
    float4 v[NV];
    uint2 braw[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane * 4 + 128 * i;
      v[i] = csrc ? load_row4(csrc + c) : load_row4(src + c);
      braw[i] = bp ? *reinterpret_cast<const uint2*>(bp + c) : make_uint2(0u, 0u);
    }
    if (bp) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 g = *reinterpret_cast<const float4*>(&sG[lane * 4 + 128 * i]);
        v[i].x += g.x * bf16_lo(braw[i].x); v[i].y += g.y * bf16_hi(braw[i].x);
        v[i].z += g.z * bf16_lo(braw[i].y); v[i].w += g.w * bf16_hi(braw[i].y);
      }
    }
    if (a.x_out) {
      XoutT* xo = reinterpret_cast<XoutT*>(a.x_out) + static_cast<long long>(in_row) * D + lane * 4;
#pragma unroll
      for (int i = 0; i < NV; ++i) store_row4(xo + 128 * i, v[i].x, v[i].y, v[i].z, v[i].w);
      if (sizeof(XoutT) == 2) {
        // the stream IS what was stored: normalise the rounded values so that forward and backward see the same x
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          v[i].x = __bfloat162float(__float2bfloat16_rn(v[i].x)); v[i].y = __bfloat162float(__float2bfloat16_rn(v[i].y));
          v[i].z = __bfloat162float(__float2bfloat16_rn(v[i].z)); v[i].w = __bfloat162float(__float2bfloat16_rn(v[i].w));
        }
      }
    }
    float s = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      s += v[i].x + v[i].y + v[i].z + v[i].w;
      s2 += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    s = warp_sum(s);
    s2 = warp_sum(s2);
    const float mean = s * (1.f / D);
    const float var = fmaxf(s2 * (1.f / D) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + 1e-6f);
    if (lane == 0) {
      if (a.mean) a.mean[r] = mean;
      if (a.rstd) a.rstd[r] = rstd;
    }
    OutT* __restrict__ op = reinterpret_cast<OutT*>(a.out) + static_cast<long long>(r) * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane * 4 + 128 * i;
      const float4 A = *reinterpret_cast<const float4*>(&sA[c]);
      const float4 B = *reinterpret_cast<const float4*>(&sB[c]);
      store_row4(op + c, fmaf((v[i].x - mean) * rstd, A.x, B.x), fmaf((v[i].y - mean) * rstd, A.y, B.y),
                 fmaf((v[i].z - mean) * rstd, A.z, B.z), fmaf((v[i].w - mean) * rstd, A.w, B.w));
    }
  }
}

template <typename OutT>
static int ln_fwd_dispatch(const LnFwdArgs& a, int D, cudaStream_t st) {
  static int target_rows = 0;
  if (target_rows == 0) {
    const char* e = getenv("UMD_LN_FWD_ROWS");
    target_rows = e ? atoi(e) : 32;
    if (target_rows < 8) target_rows = 8;
  }
  int nsamples, smax;
  if (a.gather_L > 0) {
    nsamples = a.rows_out / a.gather_L;
    smax = a.gather_L;
  } else {
    const bool second = a.rows_out > a.rm.split_row;
    nsamples = second ? a.rm.n0 + (a.rows_out - a.rm.split_row) / a.rm.s1 : a.rows_out / a.rm.s0;
    smax = a.rm.n0 > 0 ? a.rm.s0 : 0;
    if (second && a.rm.s1 > smax) smax = a.rm.s1;
  }
  if (nsamples <= 0 || smax <= 0) return UMD_OK;
  const int nchunks = ceil_div(smax, target_rows);
  const int rpc = ceil_div(smax, nchunks);
  const dim3 grid(nsamples, nchunks);
  switch (D / 128) {
#define CASE(NV)                                                                                              \
  case NV:                                                                                                    \
    if (!a.x_bf16 && !(a.x_out && a.xout_bf16)) ln_mod_fwd_kernel<NV, OutT, float, float><<<grid, 256, 0, st>>>(a, rpc);            \
    else if (!a.x_bf16) ln_mod_fwd_kernel<NV, OutT, float, __nv_bfloat16><<<grid, 256, 0, st>>>(a, rpc);                            \
    else ln_mod_fwd_kernel<NV, OutT, __nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(a, rpc);                                   \
    break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    default: set_error("ln_mod_fwd: width %d unsupported (need multiple of 128, <= 1024)", D); return UMD_ERR_UNSUPPORTED;
  }
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

int ln_mod_fwd(const LnFwdArgs& a, int D, bool out_bf16, cudaStream_t st) {
  if (a.rows_out <= 0) return UMD_OK;
  UMD_REQUIRE(D % 128 == 0 && D <= 1024, "ln_mod_fwd: width %d unsupported", D);
  UMD_REQUIRE(!(a.x_bf16 && a.x_out && !a.xout_bf16), "ln_mod_fwd: a bf16 stream cannot be written back as fp32");
  // algorithmic bytes: stream row in (+ branch), LayerNorm row out (+ stream row out)
  ProfScope prof(PC_LN_FWD, static_cast<double>(a.rows_out) * D *
                     ((a.x_bf16 ? 2 : 4) + (out_bf16 ? 2 : 4) + (a.res_branch ? 2 : 0) + (a.x_out ? (a.xout_bf16 ? 2 : 4) : 0)), st);
  return out_bf16 ? ln_fwd_dispatch<__nv_bfloat16>(a, D, st) : ln_fwd_dispatch<float>(a, D, st);
}

// =========================================================================================
// LayerNorm (+ modulate) backward.  One CTA per sample, 8 warps
// stride over the sample's rows; each lane keeps per-column partial sums for dshift, dscale,
// dgamma, dbeta which are reduced across warps through shared memory at the end.
//   dN = dY (1+scale); dscale = sum_t N dY; dshift = sum_t dY; dgamma = sum dN xhat; dbeta = sum dN
//   dx = rstd (dhat - mean(dhat) - xhat mean(dhat xhat)),  dhat = dN gamma
// =========================================================================================
// Work decomposition: grid (sample, row chunk); a row is shared by W = 1 or 2 warps (W = 2 halves the per-thread
// column state so that two CTAs fit on an SM); the per-sample sums leave by atomicAdd (several chunks per sample),
// so the caller zeroes dshift / dscale / g_dgate beforehand.
#ifndef LN_BWD_CTAS
#define LN_BWD_CTAS 2
#endif
// STAGED: a row's operands (x, dy, the dx to accumulate into, z) are brought into shared memory by per-thread cp.async
// one row ahead of their use instead of being loaded into registers when the row starts: with 126 registers per
// thread only eight rows fit into an SM's register file at once, and a warp pair had no loads in flight while it
// reduced and stored its row; the shared-memory slots keep a second row per warp pair in flight at all times.
// MODE bits: 1 = the LayerNorm input x is bf16 (residual_bf16 >= 1); 2 = the running gradient dx_in is bf16, 4 = the
// gradient written to dx is bf16 (residual_bf16 == 2: the gradient stream of the reference's bf16 flow)
template <int NV, typename DyT, bool GATE, bool STAGED, int MODE>
__global__ void __launch_bounds__(256, LN_BWD_CTAS) ln_mod_bwd_kernel(LnBwdArgs a, int rows_per_chunk) {
  constexpr bool XB = (MODE & 1) != 0, DIB = (MODE & 2) != 0, DOB = (MODE & 4) != 0;
  constexpr int D = NV * 128;
  constexpr int W = (NV % 2 == 0 && NV >= 4) ? 2 : 1;  // warps per row
  constexpr int NL = NV / W;                            // float4 columns per lane
  constexpr int RP = 8 / W;                             // rows in flight per CTA
  constexpr int CPT = (D + 255) / 256;                  // columns per thread in the final cross-warp reductions
  // sbuf: with the gate stage, the per-column sums A3 = sum_t z dx and A4 = sum_t dx of the RP rows in flight live
  // here instead of in 24 registers per thread (each element is owned by one thread: plain read-modify-write); after
  // the row loop the same memory is the scratch of the cross-warp reductions
  extern __shared__ float4 ln_bwd_dyn_smem[];           // (GATE ? 2 : 1) * RP * D floats (see ln_bwd_smem_bytes)
  float* sbuf = reinterpret_cast<float*>(ln_bwd_dyn_smem);
  __shared__ __align__(16) float s_gs[D];               // gamma * (1 + scale): d xhat = dy * s_gs
  __shared__ __align__(16) float s_gate[GATE ? D : 4];  // gate[n] (or 1): dz = gate * dx
  __shared__ float2 part[2][RP][2];
  float (*red)[D] = reinterpret_cast<float (*)[D]>(sbuf);
  float* acc3 = sbuf;
  float* acc4 = sbuf + RP * D;
  // staging slots (STAGED): per warp two slots of [x | dx | dy | z], each part lane-major (conflict-free 16 / 8 B)
  constexpr int DYB = sizeof(DyT) == 2 ? 8 : 16;
  constexpr int OFF_PV = NL * 512, OFF_DY = 2 * NL * 512, OFF_Z = OFF_DY + NL * 32 * DYB;
  constexpr int SLOT_BYTES = OFF_Z + NL * 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rp = warp / W, half = warp % W;
  const int n = blockIdx.x;
  const int S = seq_of(a.rm, n);
  const int tok_begin = blockIdx.y * rows_per_chunk;
  if (tok_begin >= S) return;
  const int tok_end = min(S, tok_begin + rows_per_chunk);
  const float* scp = a.scale ? a.scale + static_cast<long long>(n) * a.ldmod : nullptr;
  const float* gatep = (GATE && a.g_gate) ? a.g_gate + static_cast<long long>(n) * a.g_ldgate : nullptr;
  for (int c = threadIdx.x; c < D; c += 256) {
    s_gs[c] = a.gamma[c] * (scp ? 1.f + scp[c] : 1.f);
    if (GATE) s_gate[c] = gatep ? gatep[c] : 1.f;
  }
  __syncthreads();
  const int col0 = half * NL * 128 + lane * 4;
  // Every per-column sum of App. E steps 4-5 is a combination of A1 = sum_t dy and A2 = sum_t dy * xhat:
  //   dshift = A1, dscale = gamma A2 + beta A1, dbeta += (1+scale) A1, dgamma += (1+scale) A2;
  // the gate stage adds A3 = sum_t z dx and A4 = sum_t dx.
  float4 A1[NL], A2[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    A1[i] = A2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (GATE) {
      *reinterpret_cast<float4*>(&acc3[rp * D + col0 + 128 * i]) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(&acc4[rp * D + col0 + 128 * i]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  int it = 0;
  const bool want_z_s = GATE && a.g_dgate;
  const uint32_t stage_u = smem_u32(sbuf + (GATE ? 2 : 1) * RP * D) + (threadIdx.x >> 5) * 2 * SLOT_BYTES;
  auto stage_row = [&](int tok, int slot) {
    const long long off = static_cast<long long>(row_of(a.rm, n, tok)) * D + col0;
    const uint32_t sb = stage_u + slot * SLOT_BYTES;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      if (XB) cp_async8(sb + (i * 32 + lane) * 16, reinterpret_cast<const __nv_bfloat16*>(a.x) + off + 128 * i);
      else cp_async16(sb + (i * 32 + lane) * 16, reinterpret_cast<const float*>(a.x) + off + 128 * i);
      if (a.accumulate) {
        if (DIB) cp_async8(sb + OFF_PV + (i * 32 + lane) * 16, reinterpret_cast<const __nv_bfloat16*>(a.dx_in) + off + 128 * i);
        else cp_async16(sb + OFF_PV + (i * 32 + lane) * 16, reinterpret_cast<const float*>(a.dx_in) + off + 128 * i);
      }
      if (DYB == 8) cp_async8(sb + OFF_DY + (i * 32 + lane) * 8, reinterpret_cast<const DyT*>(a.dy) + off + 128 * i);
      else cp_async16(sb + OFF_DY + (i * 32 + lane) * 16, reinterpret_cast<const DyT*>(a.dy) + off + 128 * i);
      if (want_z_s) cp_async8(sb + OFF_Z + (i * 32 + lane) * 8, a.g_z + off + 128 * i);
    }
    cp_async_commit();
  };
  float mean_c = 0.f, rstd_c = 0.f;
  int slot = 0;
  if (STAGED && tok_begin + rp < tok_end) {
    stage_row(tok_begin + rp, 0);
    const int r0 = row_of(a.rm, n, tok_begin + rp);
    mean_c = a.mean[r0];
    rstd_c = a.rstd[r0];
  }
  for (int tok = tok_begin + rp; tok < tok_end; tok += RP, slot ^= 1) {
    const int xrow = row_of(a.rm, n, tok);
    const long long xoff = static_cast<long long>(xrow) * D;
    auto load_dx = [&](int c) {
      return DIB ? load_row4(reinterpret_cast<const __nv_bfloat16*>(a.dx_in) + xoff + c)
                 : load_row4(reinterpret_cast<const float*>(a.dx_in) + xoff + c);
    };
    auto store_dx = [&](int c, const float4& o) {
      if (DOB) store_row4(reinterpret_cast<__nv_bfloat16*>(a.dx) + xoff + c, o.x, o.y, o.z, o.w);
      else store_row4(reinterpret_cast<float*>(a.dx) + xoff + c, o.x, o.y, o.z, o.w);
    };
    int drow = xrow;
    bool skip = false;
    if (a.gather_L > 0) {
      const int j = tok - a.gather_off;
      skip = (j < 0 || j >= a.gather_L);
      drow = n * a.gather_L + j;
    }
    if (skip) {
      // rows outside the gathered window receive no gradient from this LayerNorm
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        const int c = col0 + 128 * i;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.accumulate) o = load_dx(c);
        if (!a.accumulate || a.dx_in != a.dx) store_dx(c, o);
        if (GATE) {
          const float4 g = *reinterpret_cast<const float4*>(&s_gate[c]);
          store_row4(a.g_dz + static_cast<long long>(xrow) * D + c, g.x * o.x, g.y * o.y, g.z * o.z, g.w * o.w);
          float4* p4 = reinterpret_cast<float4*>(&acc4[rp * D + c]);
          float4 t4 = *p4;
          t4.x += o.x; t4.y += o.y; t4.z += o.z; t4.w += o.w;
          *p4 = t4;
          if (a.g_dgate) {
            const float4 z = load_row4(a.g_z + static_cast<long long>(xrow) * D + c);
            float4* p3 = reinterpret_cast<float4*>(&acc3[rp * D + c]);
            float4 t3 = *p3;
            t3.x += z.x * o.x; t3.y += z.y * o.y; t3.z += z.z * o.z; t3.w += z.w * o.w;
            *p3 = t3;
          }
        }
      }
      continue;
    }
    float mean, rstd;
    const float* xp = reinterpret_cast<const float*>(a.x) + static_cast<long long>(xrow) * D;
    const __nv_bfloat16* xpb = reinterpret_cast<const __nv_bfloat16*>(a.x) + static_cast<long long>(xrow) * D;
    const DyT* dyp = reinterpret_cast<const DyT*>(a.dy) + static_cast<long long>(drow) * D;
    const uint32_t sb = stage_u + slot * SLOT_BYTES;
    if (STAGED) {
      // next row of this warp pair into the other slot (its previous content was consumed one iteration ago), then
      // wait for everything but that newest group: the current row has landed
      mean = mean_c;
      rstd = rstd_c;
      if (tok + RP < tok_end) {
        stage_row(tok + RP, slot ^ 1);
        const int rn = row_of(a.rm, n, tok + RP);
        mean_c = a.mean[rn];
        rstd_c = a.rstd[rn];
      } else {
        cp_async_commit();
      }
      cp_async_wait<1>();
    } else {
      mean = a.mean[drow];
      rstd = a.rstd[drow];
    }
    float4 xh[NL], dh[NL];
    float m1 = 0.f, m2 = 0.f;
    // everything this row needs from HBM is requested up front (the second half of the row's work would otherwise
    // start a second dependent round trip after the reduction)
    float4 pv[NL];
    uint2 zraw[NL];
    const bool want_z = GATE && a.g_dgate;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const int c = col0 + 128 * i;
      if (STAGED) {
        if (DIB) {
          const uint2 pr = a.accumulate ? ld_shared_v2(sb + OFF_PV + (i * 32 + lane) * 16) : make_uint2(0u, 0u);
          pv[i] = make_float4(bf16_lo(pr.x), bf16_hi(pr.x), bf16_lo(pr.y), bf16_hi(pr.y));
        } else {
          const uint4 pr = a.accumulate ? ld_shared_v4(sb + OFF_PV + (i * 32 + lane) * 16) : make_uint4(0u, 0u, 0u, 0u);
          pv[i] = make_float4(__uint_as_float(pr.x), __uint_as_float(pr.y), __uint_as_float(pr.z), __uint_as_float(pr.w));
        }
        zraw[i] = want_z ? ld_shared_v2(sb + OFF_Z + (i * 32 + lane) * 8) : make_uint2(0u, 0u);
      } else {
        pv[i] = a.accumulate ? load_dx(c) : make_float4(0.f, 0.f, 0.f, 0.f);
        zraw[i] = want_z ? *reinterpret_cast<const uint2*>(a.g_z + static_cast<long long>(xrow) * D + c) : make_uint2(0u, 0u);
      }
    }
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const int c = col0 + 128 * i;
      float4 xv, dy;
      if (STAGED) {
        if (XB) {
          const uint2 xr = ld_shared_v2(sb + (i * 32 + lane) * 16);
          xv = make_float4(bf16_lo(xr.x), bf16_hi(xr.x), bf16_lo(xr.y), bf16_hi(xr.y));
        } else {
          const uint4 xr = ld_shared_v4(sb + (i * 32 + lane) * 16);
          xv = make_float4(__uint_as_float(xr.x), __uint_as_float(xr.y), __uint_as_float(xr.z), __uint_as_float(xr.w));
        }
        if (DYB == 8) {
          const uint2 dr = ld_shared_v2(sb + OFF_DY + (i * 32 + lane) * 8);
          dy = make_float4(bf16_lo(dr.x), bf16_hi(dr.x), bf16_lo(dr.y), bf16_hi(dr.y));
        } else {
          const uint4 dr = ld_shared_v4(sb + OFF_DY + (i * 32 + lane) * 16);
          dy = make_float4(__uint_as_float(dr.x), __uint_as_float(dr.y), __uint_as_float(dr.z), __uint_as_float(dr.w));
        }
      } else {
        xv = XB ? load_row4(xpb + c) : load_row4(xp + c);
        dy = load_row4(dyp + c);
      }
      const float4 gs = *reinterpret_cast<const float4*>(&s_gs[c]);
      xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      A1[i].x += dy.x; A1[i].y += dy.y; A1[i].z += dy.z; A1[i].w += dy.w;
      A2[i].x += dy.x * xh[i].x; A2[i].y += dy.y * xh[i].y; A2[i].z += dy.z * xh[i].z; A2[i].w += dy.w * xh[i].w;
      dh[i] = make_float4(dy.x * gs.x, dy.y * gs.y, dy.z * gs.z, dy.w * gs.w);
      m1 += dh[i].x + dh[i].y + dh[i].z + dh[i].w;
      m2 += dh[i].x * xh[i].x + dh[i].y * xh[i].y + dh[i].z * xh[i].z + dh[i].w * xh[i].w;
    }
    m1 = warp_sum(m1);
    m2 = warp_sum(m2);
    if (W == 2) {
      // the two warps of a row exchange their partial sums (double-buffered slot: one named barrier per row)
      const int buf = it & 1;
      if (lane == 0) part[buf][rp][half] = make_float2(m1, m2);
      asm volatile("bar.sync %0, 64;" ::"r"(1 + rp) : "memory");
      const float2 o2 = part[buf][rp][half ^ 1];
      m1 += o2.x;
      m2 += o2.y;
      ++it;
    }
    m1 *= (1.f / D);
    m2 *= (1.f / D);
    const bool cond_tok = a.dcond && tok == 0;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const int c = col0 + 128 * i;
      float4 o = make_float4(rstd * (dh[i].x - m1 - xh[i].x * m2), rstd * (dh[i].y - m1 - xh[i].y * m2),
                             rstd * (dh[i].z - m1 - xh[i].z * m2), rstd * (dh[i].w - m1 - xh[i].w * m2));
      o.x += pv[i].x; o.y += pv[i].y; o.z += pv[i].z; o.w += pv[i].w;
      if (cond_tok) {
        float4* dc = reinterpret_cast<float4*>(a.dcond + static_cast<long long>(n) * D + c);
        float4 t = *dc;
        t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
        *dc = t;
        o = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      store_dx(c, o);
      if (GATE) {
        const float4 g = *reinterpret_cast<const float4*>(&s_gate[c]);
        store_row4(a.g_dz + static_cast<long long>(xrow) * D + c, g.x * o.x, g.y * o.y, g.z * o.z, g.w * o.w);
        float4* p4 = reinterpret_cast<float4*>(&acc4[rp * D + c]);
        float4 t4 = *p4;
        t4.x += o.x; t4.y += o.y; t4.z += o.z; t4.w += o.w;
        *p4 = t4;
        if (want_z) {
          float4* p3 = reinterpret_cast<float4*>(&acc3[rp * D + c]);
          float4 t3 = *p3;
          t3.x += bf16_lo(zraw[i].x) * o.x; t3.y += bf16_hi(zraw[i].x) * o.y;
          t3.z += bf16_lo(zraw[i].y) * o.z; t3.w += bf16_hi(zraw[i].y) * o.w;
          *p3 = t3;
        }
      }
    }
  }
  // cross-warp reduction of the column accumulators, one at a time through `red`
  auto reduce = [&](float4 (&acc)[NL], float (&out)[CPT]) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NL; ++i) *reinterpret_cast<float4*>(&red[rp][col0 + 128 * i]) = acc[i];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int c = threadIdx.x + 256 * k;
      float t = 0.f;
      if (c < D) {
#pragma unroll
        for (int w = 0; w < RP; ++w) t += red[w][c];
      }
      out[k] = t;
    }
  };
  float r1[CPT], r2[CPT];
  if (GATE) {
    // A4 / A3 are already in shared memory: sum the RP rows, then the scratch is free for A1 / A2
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int c = threadIdx.x + 256 * k;
      if (c < D) {
        float t4 = 0.f, t3 = 0.f;
#pragma unroll
        for (int w = 0; w < RP; ++w) {
          t4 += acc4[w * D + c];
          t3 += acc3[w * D + c];
        }
        if (a.g_dbias) atomicAdd(a.g_dbias + c, s_gate[c] * t4);
        if (a.g_dgate) atomicAdd(a.g_dgate + static_cast<long long>(n) * a.g_lddgate + c, t3);
      }
    }
  }
  reduce(A1, r1);
  reduce(A2, r2);
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    const int c = threadIdx.x + 256 * k;
    if (c < D) {
      const float sc1 = scp ? 1.f + scp[c] : 1.f;
      if (a.dshift) atomicAdd(a.dshift + static_cast<long long>(n) * a.ldd + c, r1[k]);
      if (a.dscale) atomicAdd(a.dscale + static_cast<long long>(n) * a.ldd + c, a.gamma[c] * r2[k] + a.beta[c] * r1[k]);
      atomicAdd(a.dgamma + c, sc1 * r2[k]);
      atomicAdd(a.dbeta + c, sc1 * r1[k]);
    }
  }
}

template <int NV, typename DyT, bool GATE, bool STAGED, int MODE>
static int ln_bwd_launch_x(const LnBwdArgs& a, dim3 grid, int rpc, int bytes, cudaStream_t st) {
  auto kern = ln_mod_bwd_kernel<NV, DyT, GATE, STAGED, MODE>;
  if (bytes > 48 * 1024) {
    static bool cfg = false;
    if (!cfg) {
      UMD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      cfg = true;
    }
  }
  kern<<<grid, 256, bytes, st>>>(a, rpc);
  return UMD_OK;
}
template <int NV, typename DyT, bool GATE, bool STAGED>
static int ln_bwd_launch_cfg(const LnBwdArgs& a, dim3 grid, int rpc, int bytes, cudaStream_t st) {
  const int mode = (a.x_bf16 ? 1 : 0) | ((a.accumulate && a.dxin_bf16) ? 2 : 0) | (a.dxout_bf16 ? 4 : 0);
  if (mode == 0) return ln_bwd_launch_x<NV, DyT, GATE, STAGED, 0>(a, grid, rpc, bytes, st);
  // the reduced-precision streams exist for the widths of the reference's variants (S 384, B 768, L 1024)
  if constexpr (NV == 3 || NV == 6 || NV == 8) {
    switch (mode) {
      case 1: return ln_bwd_launch_x<NV, DyT, GATE, STAGED, 1>(a, grid, rpc, bytes, st);   // x bf16
      case 2: return ln_bwd_launch_x<NV, DyT, GATE, STAGED, 2>(a, grid, rpc, bytes, st);   // layer 0: x fp32, dx bf16 -> fp32
      case 5: return ln_bwd_launch_x<NV, DyT, GATE, STAGED, 5>(a, grid, rpc, bytes, st);   // first of a stack: dx out bf16
      case 7: return ln_bwd_launch_x<NV, DyT, GATE, STAGED, 7>(a, grid, rpc, bytes, st);   // x, dx in, dx out bf16
      default: break;
    }
  }
  set_error("ln_mod_bwd: stream dtype combination %d is not instantiated for width %d", mode, NV * 128);
  return UMD_ERR_UNSUPPORTED;
}
template <int NV, typename DyT, bool GATE>
static int ln_bwd_launch(const LnBwdArgs& a, dim3 grid, int rpc, cudaStream_t st) {
  constexpr int W = (NV % 2 == 0 && NV >= 4) ? 2 : 1;
  constexpr int NL = NV / W;
  constexpr int BYTES = (GATE ? 2 : 1) * (8 / W) * NV * 128 * 4;
  constexpr int DYB = sizeof(DyT) == 2 ? 8 : 16;
  constexpr int STAGE_BYTES = 8 * 2 * (2 * NL * 512 + NL * 32 * DYB + NL * 256);
  static int staged_on = -1;
  if (staged_on < 0) {
    const char* e = getenv("UMD_LN_BWD_STAGED");
    staged_on = e ? atoi(e) : 1;
  }
  // the staged variant needs every row of the grid to be a LayerNorm row (no gather window); it is sized for two CTAs
  // per SM (<= 110 KB).  Width 1024 (Latent-UMD-L/2) needs 130 KB: it runs staged with one CTA per SM, which measured
  // 6.31 ms against 6.93 ms per step for the unstaged variant with two (profiles/r02_notes.md; UMD_LN_BWD_STAGED_BIG=0
  // switches back)
  static int staged_big = -1;
  if (staged_big < 0) {
    const char* e = getenv("UMD_LN_BWD_STAGED_BIG");
    staged_big = e ? atoi(e) : 1;
  }
  const int limit = staged_big ? 200 * 1024 : 110 * 1024;
  if (staged_on && a.gather_L <= 0 && BYTES + STAGE_BYTES <= limit)
    return ln_bwd_launch_cfg<NV, DyT, GATE, true>(a, grid, rpc, BYTES + STAGE_BYTES, st);
  return ln_bwd_launch_cfg<NV, DyT, GATE, false>(a, grid, rpc, BYTES, st);
}

template <typename DyT>
static int ln_bwd_dispatch(const LnBwdArgs& a, int D, int nsamples, cudaStream_t st) {
  const bool gate = a.g_dz != nullptr;
  int smax = a.rm.n0 > 0 ? a.rm.s0 : 0;
  if (nsamples > a.rm.n0 && a.rm.s1 > smax) smax = a.rm.s1;
  static int target_rows = 0;
  if (target_rows == 0) {
    const char* e = getenv("UMD_LN_BWD_ROWS");
    target_rows = e ? atoi(e) : 96;
    if (target_rows < 8) target_rows = 8;
  }
  const int nchunks = ceil_div(smax, target_rows);
  const int rpc = ceil_div(smax, nchunks);
  const dim3 grid(nsamples, nchunks);
  switch (D / 128) {
#define CASE(NV)                                                                     \
  case NV:                                                                           \
    if (gate) UMD_TRY((ln_bwd_launch<NV, DyT, true>(a, grid, rpc, st)));             \
    else UMD_TRY((ln_bwd_launch<NV, DyT, false>(a, grid, rpc, st)));                 \
    break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    default: set_error("ln_mod_bwd: width %d unsupported", D); return UMD_ERR_UNSUPPORTED;
  }
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// NOTE: dshift / dscale / g_dgate are accumulated with atomicAdd (a sample's rows are split over several CTAs):
// the caller zeroes them first.
int ln_mod_bwd(const LnBwdArgs& a, int D, int nsamples, bool dy_bf16, cudaStream_t st) {
  if (nsamples <= 0) return UMD_OK;
  UMD_REQUIRE(D % 128 == 0 && D <= 1024, "ln_mod_bwd: width %d unsupported", D);
  UMD_REQUIRE(!a.g_dgate || a.g_z, "ln_mod_bwd: the gate stage needs the saved branch output to form dgate");
  // algorithmic bytes: dy in, x in, dx read-modify-write (fp32); gate stage: dz out (bf16), z in (bf16)
  const double rows = static_cast<double>(a.rm.split_row) + static_cast<double>(nsamples - a.rm.n0) * a.rm.s1;
  ProfScope prof(PC_LN_BWD, rows * D * ((dy_bf16 ? 2 : 4) + (a.x_bf16 ? 2 : 4) + (a.accumulate ? (a.dxin_bf16 ? 2 : 4) : 0) +
                                        (a.dxout_bf16 ? 2 : 4) + (a.g_dz ? 2 : 0) + (a.g_dgate ? 2 : 0)), st);
  return dy_bf16 ? ln_bwd_dispatch<__nv_bfloat16>(a, D, nsamples, st) : ln_bwd_dispatch<float>(a, D, nsamples, st);
}

// =========================================================================================
// Gate backward (App. E steps 1 and 6): dZ = g * dX (bf16, GEMM operand), dgate[n] = sum_t Z dX,
// dbias += sum dZ.  CTA per sample; D/4 column threads x 4 row groups.
// =========================================================================================
__global__ void __launch_bounds__(1024) gate_bwd_kernel(GateBwdArgs a, int D) {
  extern __shared__ float sm[];
  const int ct = D / 4;                       // column threads
  const int col = (threadIdx.x % ct) * 4;
  const int rg = threadIdx.x / ct;            // row group 0..3
  const int n = blockIdx.x;
  const int S = seq_of(a.rm, n);
  float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
  if (a.gate) g = *reinterpret_cast<const float4*>(a.gate + static_cast<long long>(n) * a.ldgate + col);
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag;
  for (int tok = rg; tok < S; tok += 4) {
    const long long row = row_of(a.rm, n, tok);
    float4 dx = *reinterpret_cast<const float4*>(a.dx + row * D + col);
    float4 dz = make_float4(g.x * dx.x, g.y * dx.y, g.z * dx.z, g.w * dx.w);
    store_row4(a.dz + row * D + col, dz.x, dz.y, dz.z, dz.w);
    ab.x += dz.x; ab.y += dz.y; ab.z += dz.z; ab.w += dz.w;
    if (a.dgate) {
      float4 z = load_row4(a.z + row * D + col);
      ag.x += z.x * dx.x; ag.y += z.y * dx.y; ag.z += z.z * dx.z; ag.w += z.w * dx.w;
    }
  }
  float* sg = sm;            // [4][D]
  float* sb = sm + 4 * D;    // [4][D]
  *reinterpret_cast<float4*>(sg + rg * D + col) = ag;
  *reinterpret_cast<float4*>(sb + rg * D + col) = ab;
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    if (a.dgate) a.dgate[static_cast<long long>(n) * a.lddgate + c] = sg[c] + sg[D + c] + sg[2 * D + c] + sg[3 * D + c];
    if (a.dbias) atomicAdd(a.dbias + c, sb[c] + sb[D + c] + sb[2 * D + c] + sb[3 * D + c]);
  }
}

int gate_bwd(const GateBwdArgs& a, int D, int nsamples, cudaStream_t st) {
  if (nsamples <= 0) return UMD_OK;
  UMD_REQUIRE(D % 4 == 0 && D <= 1024, "gate_bwd: width %d unsupported", D);
  const double rows = static_cast<double>(a.rm.split_row) + static_cast<double>(nsamples - a.rm.n0) * a.rm.s1;
  ProfScope prof(PC_GATE_BWD, rows * D * (4 + 2 + (a.dgate ? 2 : 0)), st);  // dx in, dz out, z in
  gate_bwd_kernel<<<nsamples, D, 8 * D * sizeof(float), st>>>(a, D);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// =========================================================================================
// Column sums of a bf16 [rows, N] matrix accumulated into fp32 out[N] (bias gradients).
// No shared memory and 34 registers: the step engine runs these passes on a side stream so that they share the SMs
// with the tensor-bound GEMM that follows on the main stream (a persistent GEMM CTA leaves ~2 KB of shared memory), and
// their HBM traffic hides behind it.  Warp w of a CTA walks over rows r0 + w, r0 + w + 8, ...; a lane owns 8 adjacent
// columns and adds its partial sums with two vector reductions (red.global.add.v4.f32).
// =========================================================================================
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int rows,
                                                          int N, float* __restrict__ out, int rows_per_cta, int seg_cols,
                                                          long long seg_stride) {
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + cg * 8;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  if (col >= N) return;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int r = r0 + rl;
  // eight independent 16-byte loads in flight per thread: beside a persistent GEMM CTA only one of these CTAs fits on an
  // SM, so the bytes in flight per thread are what keeps the pass short
  for (; r + 56 < r1; r += 64) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const uint4*>(x + static_cast<long long>(r + 8 * u) * ld + col);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc[0] += bf16_lo(v[u].x); acc[1] += bf16_hi(v[u].x); acc[2] += bf16_lo(v[u].y); acc[3] += bf16_hi(v[u].y);
      acc[4] += bf16_lo(v[u].z); acc[5] += bf16_hi(v[u].z); acc[6] += bf16_lo(v[u].w); acc[7] += bf16_hi(v[u].w);
    }
  }
  for (; r < r1; r += 8) {
    uint4 v = *reinterpret_cast<const uint4*>(x + static_cast<long long>(r) * ld + col);
    acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
    acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
  }
  // column c of segment j = c / seg_cols goes to out[j * seg_stride + c % seg_cols] (q | k | v bias leaves); a lane's 8
  // columns never straddle a segment (seg_cols % 8 == 0) and every segment base is 16-byte aligned
  const long long o = seg_cols > 0 ? (col / seg_cols) * seg_stride + (col % seg_cols) : col;
  red_add_v4(out + o, acc[0], acc[1], acc[2], acc[3]);
  red_add_v4(out + o + 4, acc[4], acc[5], acc[6], acc[7]);
}

int colsum_bf16(const void* x, long long ld, int rows, int N, float* out, cudaStream_t st, int seg_cols,
                long long seg_stride) {
  if (rows <= 0) return UMD_OK;
  UMD_REQUIRE(N % 8 == 0 && ld % 8 == 0, "colsum_bf16: N and ld must be multiples of 8");
  UMD_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (seg_cols == 0 || (seg_cols % 8 == 0 && seg_stride % 4 == 0)),
              "colsum_bf16: out must be 16-byte aligned, segments multiples of 8 columns at 16-byte aligned strides");
  ProfScope prof(PC_COLSUM, static_cast<double>(rows) * N * 2, st);
  const int rpc = rows >= 8192 ? 1024 : 256;
  dim3 grid(ceil_div(N, 256), ceil_div(rows, rpc));
  colsum_bf16_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), ld, rows, N, out, rpc, seg_cols, seg_stride);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// =========================================================================================
// Model-side timestep and label vectors of one training step (see kernels.cuh)
// =========================================================================================
__global__ void step_prep_kernel(const int* __restrict__ t, const long long* __restrict__ label,
                                 const unsigned char* __restrict__ drop, int n0, int B, int num_classes, int use_labels,
                                 int* __restrict__ t_model, int* __restrict__ labels_model) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  t_model[i] = i < n0 ? t[i] + 1 : 0;
  if (labels_model) {
    int y = num_classes;
    if (use_labels && i < n0 && label && !(drop && drop[i])) y = static_cast<int>(label[i]);
    labels_model[i] = y;
  }
}
int step_prep(const int* t, const long long* label, const unsigned char* label_drop, int n0, int B, int num_classes,
              int use_labels, int* t_model, int* labels_model, cudaStream_t st) {
  if (B <= 0) return UMD_OK;
  step_prep_kernel<<<ceil_div(B, 256), 256, 0, st>>>(t, label, label_drop, n0, B, num_classes, use_labels, t_model, labels_model);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// =========================================================================================
// Conditioning path (embeddings.py:13-59, ae.py:105-124)
// =========================================================================================
// TimeEmb: emb = t * exp(-k ln(1e4)/(half-1)); out = [sin(emb) | cos(emb)] (bf16 GEMM operand).
__global__ void time_embed_kernel(const int* __restrict__ t, int B, int D, __nv_bfloat16* __restrict__ out) {
  const int half = D / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * half) return;
  const int n = idx / half, k = idx - n * half;
  const float step = logf(10000.f) / static_cast<float>(half - 1);
  const float e = static_cast<float>(t[n]) * expf(static_cast<float>(k) * -step);
  out[static_cast<long long>(n) * D + k] = __float2bfloat16(sinf(e));
  out[static_cast<long long>(n) * D + half + k] = __float2bfloat16(cosf(e));
}
int time_embed(const int* t, int B, int D, void* out, cudaStream_t st) {
  if (B <= 0) return UMD_OK;
  const int total = B * (D / 2);
  time_embed_kernel<<<ceil_div(total, 256), 256, 0, st>>>(t, B, D, reinterpret_cast<__nv_bfloat16*>(out));
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// out_bf16 = silu(h)
__global__ void silu_cast_kernel(const float* __restrict__ h, long long n, __nv_bfloat16* __restrict__ out) {
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 v = *reinterpret_cast<const float4*>(h + i);
  store_row4(out + i, silu_f(v.x), silu_f(v.y), silu_f(v.z), silu_f(v.w));
}
int silu_cast(const float* h, long long n, void* out, cudaStream_t st) {
  if (n <= 0) return UMD_OK;
  silu_cast_kernel<<<static_cast<int>(ceil_div_ll(n / 4, 256)), 256, 0, st>>>(h, n, reinterpret_cast<__nv_bfloat16*>(out));
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}
// dh_bf16 = da_bf16 * silu'(h)
__global__ void silu_bwd_kernel(const __nv_bfloat16* __restrict__ da, const float* __restrict__ h, long long n,
                                __nv_bfloat16* __restrict__ dh) {
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 v = *reinterpret_cast<const float4*>(h + i);
  float4 d = load_row4(da + i);
  store_row4(dh + i, d.x * silu_grad_f(v.x), d.y * silu_grad_f(v.y), d.z * silu_grad_f(v.z), d.w * silu_grad_f(v.w));
}
int silu_bwd(const void* da, const float* h, long long n, void* dh, cudaStream_t st) {
  if (n <= 0) return UMD_OK;
  silu_bwd_kernel<<<static_cast<int>(ceil_div_ll(n / 4, 256)), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(da), h, n, reinterpret_cast<__nv_bfloat16*>(dh));
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// s = tc + yc ; cond = adaln ? silu(s) : s   (ae.py:121-124)
__global__ void cond_combine_kernel(const float* __restrict__ tc, const float* __restrict__ yc, long long n, int adaln,
                                    float* __restrict__ s_out, float* __restrict__ cond_f32,
                                    __nv_bfloat16* __restrict__ cond_bf16) {
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 v = *reinterpret_cast<const float4*>(tc + i);
  if (yc) {
    float4 y = *reinterpret_cast<const float4*>(yc + i);
    v.x += y.x; v.y += y.y; v.z += y.z; v.w += y.w;
  }
  *reinterpret_cast<float4*>(s_out + i) = v;
  if (adaln) v = make_float4(silu_f(v.x), silu_f(v.y), silu_f(v.z), silu_f(v.w));
  *reinterpret_cast<float4*>(cond_f32 + i) = v;
  store_row4(cond_bf16 + i, v.x, v.y, v.z, v.w);
}
int cond_combine(const float* tc, const float* yc, long long n, int adaln, float* s_out, float* cond_f32,
                 void* cond_bf16, cudaStream_t st) {
  if (n <= 0) return UMD_OK;
  cond_combine_kernel<<<static_cast<int>(ceil_div_ll(n / 4, 256)), 256, 0, st>>>(
      tc, yc, n, adaln, s_out, cond_f32, reinterpret_cast<__nv_bfloat16*>(cond_bf16));
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}
// ds_bf16 = dcond * (adaln ? silu'(s) : 1)
__global__ void cond_combine_bwd_kernel(const float* __restrict__ dcond, const float* __restrict__ s, long long n,
                                        int adaln, __nv_bfloat16* __restrict__ ds) {
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 d = *reinterpret_cast<const float4*>(dcond + i);
  if (adaln) {
    float4 v = *reinterpret_cast<const float4*>(s + i);
    d.x *= silu_grad_f(v.x); d.y *= silu_grad_f(v.y); d.z *= silu_grad_f(v.z); d.w *= silu_grad_f(v.w);
  }
  store_row4(ds + i, d.x, d.y, d.z, d.w);
}
int cond_combine_bwd(const float* dcond, const float* s, long long n, int adaln, void* ds, cudaStream_t st) {
  if (n <= 0) return UMD_OK;
  cond_combine_bwd_kernel<<<static_cast<int>(ceil_div_ll(n / 4, 256)), 256, 0, st>>>(
      dcond, s, n, adaln, reinterpret_cast<__nv_bfloat16*>(ds));
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// nn.Embed lookup (embeddings.py:47): out_bf16[n] = table[ids[n]]
__global__ void gather_rows_kernel(const float* __restrict__ table, const int* __restrict__ ids, int B, int D,
                                   __nv_bfloat16* __restrict__ out) {
  const int n = blockIdx.x;
  const float* src = table + static_cast<long long>(ids[n]) * D;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(src + c);
    store_row4(out + static_cast<long long>(n) * D + c, v.x, v.y, v.z, v.w);
  }
}
int gather_rows(const float* table, const int* ids, int B, int D, void* out, cudaStream_t st) {
  if (B <= 0) return UMD_OK;
  gather_rows_kernel<<<B, 128, 0, st>>>(table, ids, B, D, reinterpret_cast<__nv_bfloat16*>(out));
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}
// transpose of the lookup: dtable[ids[n]] += d[n]  (duplicates -> atomics)
__global__ void scatter_add_rows_kernel(const float* __restrict__ d, const int* __restrict__ ids, int B, int D,
                                        float* __restrict__ dtable) {
  const int n = blockIdx.x;
  float* dst = dtable + static_cast<long long>(ids[n]) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) atomicAdd(dst + c, d[static_cast<long long>(n) * D + c]);
}
int scatter_add_rows(const float* d, const int* ids, int B, int D, float* dtable, cudaStream_t st) {
  if (B <= 0) return UMD_OK;
  scatter_add_rows_kernel<<<B, 256, 0, st>>>(d, ids, B, D, dtable);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// fp32 -> bf16
__global__ void cast_bf16_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* __restrict__ out) {
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 v = *reinterpret_cast<const float4*>(x + i);
  store_row4(out + i, v.x, v.y, v.z, v.w);
}
int cast_bf16(const float* x, long long n, void* out, cudaStream_t st) {
  if (n <= 0) return UMD_OK;
  UMD_REQUIRE(n % 4 == 0, "cast_bf16: count must be a multiple of 4");
  cast_bf16_kernel<<<static_cast<int>(ceil_div_ll(n / 4, 256)), 256, 0, st>>>(x, n, reinterpret_cast<__nv_bfloat16*>(out));
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// =========================================================================================
// q_sample (gaussian_diffusion.py:85-98): x_t = sqrt_ac[t] x0 + sqrt_1mac[t] noise
// =========================================================================================
__global__ void qsample_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const int* __restrict__ t,
                               const float* __restrict__ sqrt_ac, const float* __restrict__ sqrt_1mac, int per_sample4,
                               long long total4, float* __restrict__ out) {
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int n = static_cast<int>(i / per_sample4);
  const int tt = t[n];
  const float ca = sqrt_ac[tt], cb = sqrt_1mac[tt];
  float4 a = reinterpret_cast<const float4*>(x0)[i];
  float4 b = reinterpret_cast<const float4*>(noise)[i];
  reinterpret_cast<float4*>(out)[i] = make_float4(ca * a.x + cb * b.x, ca * a.y + cb * b.y, ca * a.z + cb * b.z, ca * a.w + cb * b.w);
}
int qsample(const float* x0, const float* noise, const int* t, const float* sqrt_ac, const float* sqrt_1mac, int n,
            int per_sample, float* out, cudaStream_t st) {
  if (n <= 0) return UMD_OK;
  UMD_REQUIRE(per_sample % 4 == 0, "q_sample: elements per sample must be a multiple of 4");
  const long long total4 = static_cast<long long>(n) * (per_sample / 4);
  qsample_kernel<<<static_cast<int>(ceil_div_ll(total4, 256)), 256, 0, st>>>(x0, noise, t, sqrt_ac, sqrt_1mac,
                                                                           per_sample / 4, total4, out);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// =========================================================================================
// DDIM update (gaussian_diffusion.py:166-284 ddim_sample on top of p_mean_variance :134-164), one fused pass:
// classifier-free-guidance combine of the doubled batch (ae.py:192-195), eps / x0 head selection
// (train_ae.py:472-483), pred_xstart (:122-127), optional clip, eps from x0 (:129-132), sigma and the Eq. 12 mean.
// x, noise, sample, pred_xstart: [n, hw, C] fp32;  pred: [n (or 2n with guidance), hw, pred_ld] fp32.
// =========================================================================================
__global__ void ddim_step_kernel(DdimArgs a, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int C = a.C;
  const long long pix = i / C;                 // n * hw + pixel
  const int c = static_cast<int>(i - pix * C);
  const int n = static_cast<int>(pix / a.hw);
  const int t = a.t[n];
  const float sra = a.sqrt_recip_ac[t], srm1 = a.sqrt_recipm1_ac[t];
  const float ab = a.ac[t];
  const float abp = a.t_next ? a.ac[a.t_next[n]] : a.ac_prev[t];
  const float x = a.x[i];
  // pred rows hold pred_ld values per pixel: [x0 head (C) | eps head (C)] (pred_ld = 2C) or the eps head alone (= C)
  const int eps_off = a.pred_ld - C;
  const long long pbase = pix * a.pred_ld + c;
  float m_eps = a.pred[pbase + eps_off];
  float m_x0 = a.eps_pred ? 0.f : a.pred[pbase];
  if (a.use_cfg) {
    const long long ubase = pbase + static_cast<long long>(a.n) * a.hw * a.pred_ld;   // unconditional half of the batch
    const float u_eps = a.pred[ubase + eps_off];
    m_eps = u_eps + a.cfg_scale * (m_eps - u_eps);
    if (!a.eps_pred) {
      const float u_x0 = a.pred[ubase];
      m_x0 = u_x0 + a.cfg_scale * (m_x0 - u_x0);
    }
  }
  const float model_eps = a.eps_pred ? m_eps : (sra * x - m_x0) / srm1;
  float px0 = sra * x - srm1 * model_eps;
  if (a.clip_denoised) px0 = fminf(fmaxf(px0, -1.f), 1.f);
  const float eps = (sra * x - px0) / srm1;
  const float sigma = a.eta * sqrtf((1.f - abp) / (1.f - ab)) * sqrtf(1.f - ab / abp);
  const float mean_pred = px0 * sqrtf(abp) + sqrtf(1.f - abp - sigma * sigma) * eps;
  const float nz = (t > 0) ? sigma * a.noise[i] : 0.f;
  a.sample[i] = mean_pred + nz;
  if (a.pred_xstart) a.pred_xstart[i] = px0;
}
int ddim_step(const DdimArgs& a, cudaStream_t st) {
  if (a.n <= 0) return UMD_OK;
  UMD_REQUIRE(a.hw > 0 && a.C > 0, "ddim_step: bad shape");
  UMD_REQUIRE(a.pred_ld == 2 * a.C || (a.pred_ld == a.C && a.eps_pred), "ddim_step: pred must hold 2C values per pixel, or C (eps head only)");
  UMD_REQUIRE(a.x && a.pred && a.noise && a.t && a.sample, "ddim_step: null argument");
  const long long total = static_cast<long long>(a.n) * a.hw * a.C;
  ddim_step_kernel<<<static_cast<int>(ceil_div_ll(total, 256)), 256, 0, st>>>(a, total);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// =========================================================================================
// random_masking index work (ae.py:14-16,25-27): stable ascending argsort of L noise values per
// row by exact rank counting (ties broken by index => identical to a stable sort), the inverse
// permutation, and the 0/1 sequence mask.  Bit-exact by construction.
// =========================================================================================
__device__ __forceinline__ uint32_t float_order_key(float f) {
  // total order: -inf < ... < -0 == +0 handled by value compare below; NaN sorts last (jnp.argsort)
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void mask_argsort_kernel(const float* __restrict__ noise, int L, int keep, int* __restrict__ ids_shuffle,
                                    int* __restrict__ ids_restore, float* __restrict__ mask) {
  extern __shared__ uint32_t keys[];
  const int n = blockIdx.x;
  const float* row = noise + static_cast<long long>(n) * L;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    float f = row[i];
    uint32_t k;
    if (f != f) k = 0xFFFFFFFFu;              // NaN last
    else if (f == 0.f) k = 0x80000000u;       // -0 and +0 compare equal
    else k = float_order_key(f);
    keys[i] = k;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const uint32_t ki = keys[i];
    int rank = 0;
    for (int j = 0; j < L; ++j) {
      const uint32_t kj = keys[j];
      rank += (kj < ki) || (kj == ki && j < i);
    }
    ids_shuffle[static_cast<long long>(n) * L + rank] = i;
    ids_restore[static_cast<long long>(n) * L + i] = rank;
    if (mask) mask[static_cast<long long>(n) * L + i] = rank >= keep ? 1.f : 0.f;
  }
}
int mask_argsort(const float* noise, int n, int L, int keep, int* ids_shuffle, int* ids_restore, float* mask,
                 cudaStream_t st) {
  if (n <= 0) return UMD_OK;
  UMD_REQUIRE(L > 0 && L <= 8192, "mask_argsort: L=%d unsupported", L);
  mask_argsort_kernel<<<n, 256, L * sizeof(uint32_t), st>>>(noise, L, keep, ids_shuffle, ids_restore, mask);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// =========================================================================================
// Patch embedding, gather-then-embed (ae.py:64-66,101-103,130,19-22,139): only the kept patches
// are embedded; result is identical to embed-all-then-gather.  x_enc row (n, tok0+num_cls+k) =
// W^T patch(ids_keep[n,k]) + b + pos[ids_keep[n,k]];  rows (n, tok0+c) = cls[c].
// =========================================================================================
template <int TOK>
__global__ void __launch_bounds__(256) embed_fwd_kernel(EmbedArgs a) {
  __shared__ float patch[TOK][64];
  __shared__ int pid_s[TOK];
  const int n = blockIdx.y;
  const int seg = n < a.rm.n0 ? 0 : 1;
  const int keep = seg ? a.keep1 : a.keep0;
  const int masked = seg ? a.masked1 : a.masked0;
  const int k0 = blockIdx.x * TOK;
  if (k0 >= keep) return;
  const int PK = a.patch * a.patch * a.C;
  const int gw = a.img / a.patch;
  for (int i = threadIdx.x; i < TOK * PK; i += blockDim.x) {
    const int tk = i / PK, e = i - tk * PK;
    const int k = k0 + tk;
    float v = 0.f;
    if (k < keep) {
      const int pid = masked ? a.ids_keep[static_cast<long long>(n) * a.L + k] : k;
      if (e == 0) pid_s[tk] = pid;
      const int c = e % a.C, ab = e / a.C;
      const int pa = ab / a.patch, pb = ab - pa * a.patch;
      const int py = (pid / gw) * a.patch + pa, px = (pid % gw) * a.patch + pb;
      v = a.image[((static_cast<long long>(n) * a.img + py) * a.img + px) * a.C + c];
    }
    patch[tk][e] = v;
  }
  __syncthreads();
  for (int col = threadIdx.x * 4; col < a.D; col += blockDim.x * 4) {
    float4 acc[TOK];
#pragma unroll
    for (int tk = 0; tk < TOK; ++tk) acc[tk] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = 0; e < PK; ++e) {
      const float4 w = *reinterpret_cast<const float4*>(a.W + static_cast<long long>(e) * a.D + col);
#pragma unroll
      for (int tk = 0; tk < TOK; ++tk) {
        const float pv = patch[tk][e];
        acc[tk].x += pv * w.x; acc[tk].y += pv * w.y; acc[tk].z += pv * w.z; acc[tk].w += pv * w.w;
      }
    }
    const float4 b = *reinterpret_cast<const float4*>(a.bias + col);
#pragma unroll
    for (int tk = 0; tk < TOK; ++tk) {
      const int k = k0 + tk;
      if (k < keep) {
        const float4 p = *reinterpret_cast<const float4*>(a.pos + static_cast<long long>(pid_s[tk]) * a.D + col);
        const long long row = row_of(a.rm, n, a.tok0 + a.num_cls + k);
        *reinterpret_cast<float4*>(a.x + row * a.D + col) =
            make_float4(acc[tk].x + b.x + p.x, acc[tk].y + b.y + p.y, acc[tk].z + b.z + p.z, acc[tk].w + b.w + p.w);
      }
    }
  }
}
__global__ void fill_cls_kernel(EmbedArgs a) {
  const int n = blockIdx.x, c = blockIdx.y;
  const long long row = row_of(a.rm, n, a.tok0 + c);
  for (int col = threadIdx.x * 4; col < a.D; col += blockDim.x * 4)
    *reinterpret_cast<float4*>(a.x + row * a.D + col) = *reinterpret_cast<const float4*>(a.cls + static_cast<long long>(c) * a.D + col);
}
int embed_fwd(const EmbedArgs& a, int nsamples, cudaStream_t st) {
  if (nsamples <= 0) return UMD_OK;
  UMD_REQUIRE(a.patch * a.patch * a.C <= 64, "embed_fwd: patch vector longer than 64 is not supported");
  UMD_REQUIRE(a.D % 4 == 0, "embed_fwd: width must be a multiple of 4");
  const int maxkeep = a.keep0 > a.keep1 ? a.keep0 : a.keep1;
  dim3 grid(ceil_div(maxkeep, 8), nsamples);
  embed_fwd_kernel<8><<<grid, 256, 0, st>>>(a);
  UMD_LAUNCH_CHECK();
  fill_cls_kernel<<<dim3(nsamples, a.num_cls), 256, 0, st>>>(a);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// Backward of the embedding: dW[e,:] += sum patch[e] dx, db += sum dx, dpos[pid] += dx, dcls[c] += dx.
template <int KE>
__global__ void __launch_bounds__(256) embed_bwd_kernel(EmbedArgs a, const float* __restrict__ dx, float* __restrict__ dW,
                                                        float* __restrict__ db, float* __restrict__ dpos, int rows_per_cta) {
  // grid.x: chunk of kept-token rows over the flattened (sample, k) space; grid.y: slice of KE patch elements.
  // Thread t owns the float4 column group t (D <= 1024 => at most 256 groups).
  __shared__ float pe[64][KE];   // [row-in-chunk][e]
  __shared__ long long xrow_s[64];
  __shared__ int pid_s[64];
  const int PK = a.patch * a.patch * a.C;
  const int e0 = blockIdx.y * KE;
  const int gw = a.img / a.patch;
  const long long total0 = static_cast<long long>(a.rm.n0) * a.keep0;
  const long long total = total0 + static_cast<long long>(a.n1) * a.keep1;
  const long long base = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long lim = min(total, base + rows_per_cta);
  const int cg = threadIdx.x;
  const bool active = cg < a.D / 4;
  float4 acc[KE];
#pragma unroll
  for (int e = 0; e < KE; ++e) acc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 accb = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r0 = base; r0 < lim; r0 += 64) {
    __syncthreads();
    const long long rem = lim - r0;
    const int nr = rem < 64 ? static_cast<int>(rem) : 64;
    for (int i = threadIdx.x; i < 64 * KE; i += 256) {
      const int rr = i / KE, e = i - rr * KE;
      float v = 0.f;
      if (rr < nr) {
        const long long f = r0 + rr;
        int n, k, masked;
        if (f < total0) {
          n = static_cast<int>(f / a.keep0);
          k = static_cast<int>(f - static_cast<long long>(n) * a.keep0);
          masked = a.masked0;
        } else {
          const long long g = f - total0;
          const int q = static_cast<int>(g / a.keep1);
          n = a.rm.n0 + q;
          k = static_cast<int>(g - static_cast<long long>(q) * a.keep1);
          masked = a.masked1;
        }
        const int pid = masked ? a.ids_keep[static_cast<long long>(n) * a.L + k] : k;
        if (e == 0) {
          pid_s[rr] = pid;
          xrow_s[rr] = row_of(a.rm, n, a.tok0 + a.num_cls + k);
        }
        const int ee = e0 + e;
        if (ee < PK) {
          const int c = ee % a.C, ab = ee / a.C;
          const int pa = ab / a.patch, pb = ab - pa * a.patch;
          const int py = (pid / gw) * a.patch + pa, px = (pid % gw) * a.patch + pb;
          v = a.image[((static_cast<long long>(n) * a.img + py) * a.img + px) * a.C + c];
        }
      }
      pe[rr][e] = v;
    }
    __syncthreads();
    if (active) {
      // four rows of dx in flight per thread (one dependent load per iteration left the kernel at ~1.2 TB/s)
      for (int rr = 0; rr < nr; rr += 4) {
        float4 d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          d[u] = (rr + u < nr) ? *reinterpret_cast<const float4*>(dx + xrow_s[rr + u] * a.D + cg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (rr + u < nr) {
#pragma unroll
            for (int e = 0; e < KE; ++e) {
              const float pv = pe[rr + u][e];
              acc[e].x += pv * d[u].x; acc[e].y += pv * d[u].y; acc[e].z += pv * d[u].z; acc[e].w += pv * d[u].w;
            }
            if (blockIdx.y == 0) {
              accb.x += d[u].x; accb.y += d[u].y; accb.z += d[u].z; accb.w += d[u].w;
              red_add_v4(dpos + static_cast<long long>(pid_s[rr + u]) * a.D + cg * 4, d[u].x, d[u].y, d[u].z, d[u].w);
            }
          }
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int e = 0; e < KE; ++e) {
      if (e0 + e < PK) red_add_v4(dW + static_cast<long long>(e0 + e) * a.D + cg * 4, acc[e].x, acc[e].y, acc[e].z, acc[e].w);
    }
    if (blockIdx.y == 0) red_add_v4(db + cg * 4, accb.x, accb.y, accb.z, accb.w);
  }
}
__global__ void cls_bwd_kernel(EmbedArgs a, const float* __restrict__ dx, float* __restrict__ dcls, int nsamples) {
  const int c = blockIdx.x;
  for (int col = threadIdx.x; col < a.D; col += blockDim.x) {
    float t = 0.f;
    for (int n = blockIdx.y; n < nsamples; n += gridDim.y) t += dx[static_cast<long long>(row_of(a.rm, n, a.tok0 + c)) * a.D + col];
    atomicAdd(dcls + static_cast<long long>(c) * a.D + col, t);
  }
}
int embed_bwd(const EmbedArgs& a, int nsamples, const float* dx, float* dW, float* db, float* dpos, float* dcls,
              cudaStream_t st) {
  if (nsamples <= 0) return UMD_OK;
  const int PK = a.patch * a.patch * a.C;
  const long long total = static_cast<long long>(a.rm.n0) * a.keep0 + static_cast<long long>(a.n1) * a.keep1;
  const int rpc = 256;
  dim3 grid(static_cast<int>(ceil_div_ll(total, rpc)), ceil_div(PK, 16));
  embed_bwd_kernel<16><<<grid, 256, 0, st>>>(a, dx, dW, db, dpos, rpc);
  UMD_LAUNCH_CHECK();
  cls_bwd_kernel<<<dim3(a.num_cls, 16), 256, 0, st>>>(a, dx, dcls, nsamples);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// x'[n, 0, :] = cond[n]  (adaln=False: conditioning token prepended inside every block, vit.py:73-74)
__global__ void set_cond_row_kernel(float* __restrict__ x, const float* __restrict__ cond, RowMap rm, int D) {
  const int n = blockIdx.x;
  const long long row = row_of(rm, n, 0);
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4)
    *reinterpret_cast<float4*>(x + row * D + c) = *reinterpret_cast<const float4*>(cond + static_cast<long long>(n) * D + c);
}
int set_cond_row(float* x, const float* cond, const RowMap& rm, int nsamples, int D, cudaStream_t st) {
  if (nsamples <= 0) return UMD_OK;
  set_cond_row_kernel<<<nsamples, 256, 0, st>>>(x, cond, rm, D);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}
// transpose: dcond[n] += dx'[n,0]; dx'[n,0] = 0 (the block's output row 0 is discarded, vit.py:111-112)
__global__ void cond_row_bwd_kernel(float* __restrict__ dx, float* __restrict__ dcond, RowMap rm, int D) {
  const int n = blockIdx.x;
  const long long row = row_of(rm, n, 0);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    dcond[static_cast<long long>(n) * D + c] += dx[row * D + c];
    dx[row * D + c] = 0.f;
  }
}
int cond_row_bwd(float* dx, float* dcond, const RowMap& rm, int nsamples, int D, cudaStream_t st) {
  if (nsamples <= 0) return UMD_OK;
  cond_row_bwd_kernel<<<nsamples, 256, 0, st>>>(dx, dcond, rm, D);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// =========================================================================================
// Decoder input (ae.py:142,147-162): row 0 = rep = mean of the num_cls encoder outputs; row 1+j =
// (rank[j] < keep ? enc[rank[j]] : mask_token) + dec_pos[j].
// =========================================================================================
__global__ void decoder_input_fwd_kernel(DecInArgs a) {
  const int n = blockIdx.y, r = blockIdx.x;  // r in [0, 1+L)
  const int seg = n < a.rm_enc.n0 ? 0 : 1;
  const int keep = seg ? a.keep1 : a.keep0;
  const int masked = seg ? a.masked1 : a.masked0;
  const long long orow = (static_cast<long long>(n) * a.S_d + a.tok0 + r) * a.D;
  if (r == 0) {
    const float inv = 1.f / a.num_cls;
    for (int c = threadIdx.x * 4; c < a.D; c += blockDim.x * 4) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < a.num_cls; ++q) {
        float4 v = *reinterpret_cast<const float4*>(a.enc + static_cast<long long>(row_of(a.rm_enc, n, a.tok0 + q)) * a.D + c);
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      t.x *= inv; t.y *= inv; t.z *= inv; t.w *= inv;
      *reinterpret_cast<float4*>(a.xd + orow + c) = t;
      if (a.rep) *reinterpret_cast<float4*>(a.rep + static_cast<long long>(n) * a.D + c) = t;
    }
    return;
  }
  const int j = r - 1;
  const int rank = masked ? a.ids_restore[static_cast<long long>(n) * a.L + j] : j;
  const float* src = rank < keep ? a.enc + static_cast<long long>(row_of(a.rm_enc, n, a.tok0 + a.num_cls + rank)) * a.D
                                 : a.mask_token;
  const float* pos = a.dec_pos + static_cast<long long>(j) * a.D;
  for (int c = threadIdx.x * 4; c < a.D; c += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(src + c);
    float4 p = *reinterpret_cast<const float4*>(pos + c);
    *reinterpret_cast<float4*>(a.xd + orow + c) = make_float4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
  }
}
int decoder_input_fwd(const DecInArgs& a, int nsamples, cudaStream_t st) {
  if (nsamples <= 0) return UMD_OK;
  decoder_input_fwd_kernel<<<dim3(1 + a.L, nsamples), 192, 0, st>>>(a);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// Backward part 1: gradient w.r.t. the encoder's final-LN output (every encoder row written once).
__global__ void decoder_input_bwd_enc_kernel(DecInArgs a, const float* __restrict__ dxd, float* __restrict__ denc) {
  const int row = blockIdx.x;
  const int n = sample_of(a.rm_enc, row);
  const int tok = token_of(a.rm_enc, row);
  const int seg = n < a.rm_enc.n0 ? 0 : 1;
  const int masked = seg ? a.masked1 : a.masked0;
  float scale = 1.f;
  const float* src = nullptr;
  if (tok < a.tok0) {
    src = nullptr;
  } else if (tok < a.tok0 + a.num_cls) {
    src = dxd + (static_cast<long long>(n) * a.S_d + a.tok0) * a.D;
    scale = 1.f / a.num_cls;
  } else {
    const int k = tok - a.tok0 - a.num_cls;
    const int pid = masked ? a.ids_keep[static_cast<long long>(n) * a.L + k] : k;
    src = dxd + (static_cast<long long>(n) * a.S_d + a.tok0 + 1 + pid) * a.D;
  }
  for (int c = threadIdx.x * 4; c < a.D; c += blockDim.x * 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src) {
      v = *reinterpret_cast<const float4*>(src + c);
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    }
    *reinterpret_cast<float4*>(denc + static_cast<long long>(row) * a.D + c) = v;
  }
}
// Backward part 2: ddec_pos[j] += sum_n dxd[n,1+j];  dmask_token += sum over masked (n,j).
__global__ void decoder_input_bwd_param_kernel(DecInArgs a, const float* __restrict__ dxd, float* __restrict__ ddec_pos,
                                               float* __restrict__ dmask_token, int nsamples) {
  const int j = blockIdx.x;
  for (int c = threadIdx.x * 4; c < a.D; c += blockDim.x * 4) {
    float4 tp = make_float4(0.f, 0.f, 0.f, 0.f), tm = tp;
    for (int n = blockIdx.y; n < nsamples; n += gridDim.y) {
      const int seg = n < a.rm_enc.n0 ? 0 : 1;
      const int keep = seg ? a.keep1 : a.keep0;
      const int masked = seg ? a.masked1 : a.masked0;
      float4 v = *reinterpret_cast<const float4*>(dxd + (static_cast<long long>(n) * a.S_d + a.tok0 + 1 + j) * a.D + c);
      tp.x += v.x; tp.y += v.y; tp.z += v.z; tp.w += v.w;
      if (masked && a.ids_restore[static_cast<long long>(n) * a.L + j] >= keep) {
        tm.x += v.x; tm.y += v.y; tm.z += v.z; tm.w += v.w;
      }
    }
    red_add_v4(ddec_pos + static_cast<long long>(j) * a.D + c, tp.x, tp.y, tp.z, tp.w);
    red_add_v4(dmask_token + c, tm.x, tm.y, tm.z, tm.w);
  }
}
int decoder_input_bwd(const DecInArgs& a, int nsamples, int enc_rows, const float* dxd, float* denc, float* ddec_pos,
                      float* dmask_token, cudaStream_t st) {
  if (nsamples <= 0) return UMD_OK;
  decoder_input_bwd_enc_kernel<<<enc_rows, 192, 0, st>>>(a, dxd, denc);
  UMD_LAUNCH_CHECK();
  decoder_input_bwd_param_kernel<<<dim3(a.L, 8), 192, 0, st>>>(a, dxd, ddec_pos, dmask_token, nsamples);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// =========================================================================================
// Loss (train_ae.py:333-360) fused with its gradient (App. E "Loss -> dpred").
// predp is the un-patchify GEMM output in patch-major layout [n*L, p*p*2C] (column = (a*p+b)*2C+o,
// already oriented so that (a,b) is the pixel offset inside the patch — see final_conv packing).
// One thread per (row, a, b): reads 2C predictions, C x0 targets, C noise targets.
// Partial sums per CTA -> loss_finalize.
// =========================================================================================
__global__ void __launch_bounds__(256) loss_kernel(LossArgs a) {
  __shared__ float sred[8];
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int pp = a.patch * a.patch;
  const long long total = static_cast<long long>(a.n0 + a.n1) * a.L * pp;
  float part = 0.f;
  if (idx < total) {
    const long long row = idx / pp;
    const int ab = static_cast<int>(idx - row * pp);
    const int n = static_cast<int>(row / a.L), l = static_cast<int>(row - static_cast<long long>(n) * a.L);
    const int seg = n < a.n0 ? 0 : 1;
    const int keep = seg ? a.keep1 : a.keep0;
    const int masked = seg ? a.masked1 : a.masked0;
    const float m = masked ? (a.ids_restore[static_cast<long long>(n) * a.L + l] >= keep ? 1.f : 0.f) : 1.f;
    // weights of the squared errors for this pixel
    float wx, we;
    if (seg == 0) {
      if (masked) { wx = a.w_x0_0 * m; we = a.w_eps_0 * (1.f - m); }
      else { wx = a.w_x0_0; we = a.w_eps_0; }
    } else {
      wx = a.w_x0_1 * m; we = 0.f;
    }
    const int gw = a.img / a.patch;
    const int pa = ab / a.patch, pb = ab - pa * a.patch;
    const int py = (l / gw) * a.patch + pa, px = (l % gw) * a.patch + pb;
    const long long pix = ((static_cast<long long>(n) * a.img + py) * a.img + px) * a.C;
    const int NC = pp * 2 * a.C;
    const float* pr = a.predp + row * NC + ab * 2 * a.C;
    __nv_bfloat16* dp = a.dpredp ? a.dpredp + row * NC + ab * 2 * a.C : nullptr;
    for (int o = 0; o < a.C; ++o) {
      const float d = pr[o] - a.x0[pix + o];
      part += wx * d * d;
      if (dp) dp[o] = __float2bfloat16(2.f * wx * d * a.grad_scale);
    }
    for (int o = 0; o < a.C; ++o) {
      float d = 0.f;
      if (seg == 0) d = pr[a.C + o] - a.noise[pix + o];
      part += we * d * d;
      if (dp) dp[a.C + o] = __float2bfloat16(2.f * we * d * a.grad_scale);
    }
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = sred[threadIdx.x];
    v += __shfl_xor_sync(0xffu, v, 4);
    v += __shfl_xor_sync(0xffu, v, 2);
    v += __shfl_xor_sync(0xffu, v, 1);
    if (threadIdx.x == 0) a.partials[blockIdx.x] = v;
  }
}
__global__ void loss_finalize_kernel(const float* __restrict__ partials, int n, float* __restrict__ loss_out) {
  __shared__ double sm[256];
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) t += static_cast<double>(partials[i]);
  sm[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = static_cast<float>(sm[0]);
}
int loss_num_partials(const LossArgs& a) {
  const long long total = static_cast<long long>(a.n0 + a.n1) * a.L * a.patch * a.patch;
  return static_cast<int>(ceil_div_ll(total, 256));
}
int loss_fwd_bwd(const LossArgs& a, float* loss_out, cudaStream_t st) {
  const int np = loss_num_partials(a);
  if (np <= 0) return UMD_OK;
  loss_kernel<<<np, 256, 0, st>>>(a);
  UMD_LAUNCH_CHECK();
  loss_finalize_kernel<<<1, 256, 0, st>>>(a.partials, np, loss_out);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// pred[n, i*p+a, j*p+b, o] = predp[(n,l=(i,j)), (a*p+b)*2C+o]   (ae.py:172-173 output layout)
__global__ void unpatchify_kernel(const float* __restrict__ predp, int n_total, int img, int patch, int C2,
                                  float* __restrict__ pred) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n_total) * img * img * C2;
  if (idx >= total) return;
  const int o = static_cast<int>(idx % C2);
  long long r = idx / C2;
  const int px = static_cast<int>(r % img); r /= img;
  const int py = static_cast<int>(r % img);
  const int n = static_cast<int>(r / img);
  const int gw = img / patch;
  const int l = (py / patch) * gw + px / patch;
  const int ab = (py % patch) * patch + px % patch;
  pred[idx] = predp[(static_cast<long long>(n) * gw * gw + l) * (patch * patch * C2) + ab * C2 + o];
}
int unpatchify(const float* predp, int n, int img, int patch, int C2, float* pred, cudaStream_t st) {
  if (n <= 0) return UMD_OK;
  const long long total = static_cast<long long>(n) * img * img * C2;
  unpatchify_kernel<<<static_cast<int>(ceil_div_ll(total, 256)), 256, 0, st>>>(predp, n, img, patch, C2, pred);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// =========================================================================================
// final_conv weight packing (ae.py:95-97; App. A.7).  Flax kernel K[p,p,D,2C] -> bf16 matrix
// Wm[c, (a*p+b)*2C+o] = K[a', b', c, o] with (a',b') = flip ? (p-1-a, p-1-b) : (a,b);
// biasm[(a*p+b)*2C+o] = bias[o].  The transpose routine adds the matrix-layout gradient back.
// =========================================================================================
__global__ void pack_final_conv_kernel(const float* __restrict__ K, const float* __restrict__ bias, int p, int D, int C2,
                                       int flip, __nv_bfloat16* __restrict__ Wm, float* __restrict__ biasm) {
  const int NC = p * p * C2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < D * NC) {
    const int c = idx / NC, col = idx - c * NC;
    const int o = col % C2, ab = col / C2;
    int pa = ab / p, pb = ab - pa * p;
    if (flip) { pa = p - 1 - pa; pb = p - 1 - pb; }
    Wm[idx] = __float2bfloat16(K[((static_cast<long long>(pa) * p + pb) * D + c) * C2 + o]);
  }
  if (idx < NC) biasm[idx] = bias[idx % C2];
}
int pack_final_conv(const float* K, const float* bias, int p, int D, int C2, int flip, void* Wm, float* biasm,
                    cudaStream_t st) {
  const int total = D * p * p * C2;
  pack_final_conv_kernel<<<ceil_div(total, 256), 256, 0, st>>>(K, bias, p, D, C2, flip, reinterpret_cast<__nv_bfloat16*>(Wm), biasm);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}
__global__ void unpack_final_conv_grad_kernel(const float* __restrict__ dWm, const float* __restrict__ dbiasm, int p,
                                              int D, int C2, int flip, float* __restrict__ dK, float* __restrict__ dbias) {
  const int NC = p * p * C2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < D * NC) {
    const int c = idx / NC, col = idx - c * NC;
    const int o = col % C2, ab = col / C2;
    int pa = ab / p, pb = ab - pa * p;
    if (flip) { pa = p - 1 - pa; pb = p - 1 - pb; }
    dK[((static_cast<long long>(pa) * p + pb) * D + c) * C2 + o] += dWm[idx];
  }
  if (idx < C2) {
    float t = 0.f;
    for (int ab = 0; ab < p * p; ++ab) t += dbiasm[ab * C2 + idx];
    dbias[idx] += t;
  }
}
int unpack_final_conv_grad(const float* dWm, const float* dbiasm, int p, int D, int C2, int flip, float* dK, float* dbias,
                           cudaStream_t st) {
  const int total = D * p * p * C2;
  unpack_final_conv_grad_kernel<<<ceil_div(total, 256), 256, 0, st>>>(dWm, dbiasm, p, D, C2, flip, dK, dbias);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

// sum over samples of a [B, N] fp32 matrix into out[N] (+=): bias gradient of the adaLN projections
__global__ void colsum_f32_kernel(const float* __restrict__ x, long long ld, int rows, int N, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float t = 0.f;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) t += x[static_cast<long long>(r) * ld + c];
  atomicAdd(out + c, t);
}
int colsum_f32(const float* x, long long ld, int rows, int N, float* out, cudaStream_t st) {
  if (rows <= 0) return UMD_OK;
  colsum_f32_kernel<<<dim3(ceil_div(N, 256), 8), 256, 0, st>>>(x, ld, rows, N, out);
  UMD_LAUNCH_CHECK();
  return UMD_OK;
}

}  // namespace umd
