// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring ->
// tcgen05.mma (single issuing thread, fp32 accumulators in TMEM, double-buffered) ->
// tcgen05.ld epilogue with the fused element-wise tails the UMD block needs.
//
// Covers every dense contraction on the training-step path of the reference
// (big_vision/models/vit.py:54,57 MLP; :71 adaLN projection; :82-87 q/k/v/out projections;
//  big_vision/models/ae.py:94-97 final modulation / un-patchify; embeddings.py:56,58 trunks)
// and their autodiff transposes (dgrad: A K-major x W K-major; wgrad: both MN-major).
//
// One CTA per SM, 384 threads:
//   warp 0 lane 0 : TMA producer           warp 1 lane 0 : MMA issuer
//   warp 2        : TMEM allocator         warps 4..11   : epilogue (2 warps per TMEM lane quadrant)
#include "common.cuh"
#include "ptx.cuh"

#include <stdlib.h>
#include <string.h>

namespace umd {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 384;

struct GemmParams {
  int M, N, K, batch, split_k;
  int a_bcast, b_bcast, b_kchunk;
  int cta2;   // run on CTA pairs (256-row tiles)
  void* out0;
  long long ld0, bs0;
  void* out1;
  long long ld1;
  const float* bias;
  long long bias_bs;
  const void* aux;
  long long ldaux;
  const float* gate;
  long long ldgate;
  RowMap rmap;
};

// CTA2: the tile is 256 x BN over a pair of CTAs (tcgen05 cta_group::2).  Each CTA stages its 128 rows of A and
// half of B's BN rows, so a stage is 32 KB instead of 48 KB and the ring is 5 deep (~2500 cycles of look-ahead
// instead of ~2000, plus room for double-buffered epilogue staging), which is what the K = 768 GEMMs of the block need to stop waiting for TMA.
template <int BN, bool CTA2 = false>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (CTA2 ? BN / 2 : BN) * BK * 2;
  static constexpr int STAGES = CTA2 ? 5 : ((BN == 256) ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int OUT_BUFS = CTA2 ? 2 : 1;   // staging chunks per epilogue warp (2: a store drains while the next is staged)
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  // epilogue staging: one [32 rows][64 bf16] swizzled chunk per epilogue warp, drained by TMA stores
  static constexpr int STAGE_OUT_BYTES = (BN >= 128) ? 8 * OUT_BUFS * 4096 : 0;
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + STAGE_OUT_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
};

// nn.gelu (tanh form, vit.py:55) and its derivative, written as explicit FMA chains: the fc1 / fc2-dgrad GEMMs have
// K = 768 only, so their epilogues (these functions once per output element) must stay well under the main loop's
// ~3000 cycles per 256 x 256 tile to hide behind it.  5 FP + 1 MUFU and 10 FP + 1 MUFU per element.
__device__ __forceinline__ float gelu_tanh(float u) {
  const float k0 = 0.7978845608028654f, k0k1 = 0.7978845608028654f * 0.044715f;
  const float u2 = u * u;
  const float t = tanh_fast(u * fmaf(u2, k0k1, k0));
  const float hu = 0.5f * u;
  return fmaf(hu, t, hu);
}
__device__ __forceinline__ float gelu_tanh_grad(float u) {
  const float k0 = 0.7978845608028654f, k0k1 = 0.7978845608028654f * 0.044715f;
  const float u2 = u * u;
  const float t = tanh_fast(u * fmaf(u2, k0k1, k0));
  const float q = fmaf(u2, 3.f * k0k1, k0);        // k0 (1 + 3 k1 u^2)
  const float s = fmaf(-t, t, 1.f);                // 1 - tanh^2
  const float w = (0.5f * u) * s;
  return fmaf(w, q, fmaf(0.5f, t, 0.5f));
}

struct WorkItem {
  int m0, n0, b, kb0, kb1;
};
// item order: n fastest, then batch, then m, then k-split.  CTAs that run concurrently then share the A row-panel (all n
// and batch items of one m) AND, in a split-K weight gradient, the B column-panel (all m of one k-range): with the k-split
// inside m the three m-tiles of a K = 131 584 wgrad ran in different waves and re-read the B panel from HBM once each
// (ncu: 2.0 GB against 0.8 GB algorithmic for the batched q/k/v weight gradient).
template <int BN, int TM = BM>
__device__ __forceinline__ WorkItem decode_item(long long item, int n_tiles, int m_tiles, int batch, int split_k, int kb_total) {
  WorkItem w;
  const int n = static_cast<int>(item % n_tiles);
  long long r = item / n_tiles;
  w.b = static_cast<int>(r % batch);
  r /= batch;
  const int m = static_cast<int>(r % m_tiles);
  const int ks = static_cast<int>(r / m_tiles);
  w.m0 = m * TM;
  w.n0 = n * BN;
  w.kb0 = static_cast<int>(static_cast<long long>(ks) * kb_total / split_k);
  w.kb1 = static_cast<int>(static_cast<long long>(ks + 1) * kb_total / split_k);
  return w;
}

template <int BN, bool A_MN, bool B_MN, int EPI, bool CTA2>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
            const __grid_constant__ CUtensorMap tmAux, const GemmParams p) {
  using Cfg = GemmCfg<BN, CTA2>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int TM = CTA2 ? 2 * BM : BM;   // rows of one work item
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* smem_out = smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + Cfg::STAGE_OUT_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint64_t* aux_bar = bars + 2 * STAGES + 4;      // [8] one per epilogue warp (DGELU operand tiles)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 12);
  // bf16 results leave through swizzled shared memory and TMA stores (full 128-byte lines, M/N tails clipped by
  // the tensor map) instead of one 16-byte store per lane per row
  constexpr bool STAGED = (BN >= 128) && (EPI == UMD_EPI_BF16 || EPI == UMD_EPI_GELU || EPI == UMD_EPI_DGELU || EPI == UMD_EPI_BF16_DELTA);
  constexpr bool AUX_TILE = (EPI == UMD_EPI_DGELU || EPI == UMD_EPI_BF16_DELTA);   // the epilogue reads a bf16 operand tile fetched by TMA

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CTA2 ? static_cast<int>(cluster_ctarank()) : 0;   // 0 = leader (issues the MMAs of the pair)
  const int worker = CTA2 ? blockIdx.x >> 1 : blockIdx.x;             // work items are dealt to CTA pairs
  const int nworkers = CTA2 ? gridDim.x >> 1 : gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (STAGED) {
      tma_prefetch_desc(&tmO0);
      if (EPI == UMD_EPI_GELU) tma_prefetch_desc(&tmO1);
      if (AUX_TILE) tma_prefetch_desc(&tmAux);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], CTA2 ? 16 : 8);   // the leader's copy collects the epilogue warps of both CTAs
    }
    for (int i = 0; i < 8; ++i) mbar_init(&aux_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CTA2) {
      tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();   // the pair's barriers exist before either CTA signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = (p.M + TM - 1) / TM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int kb_total = (p.K + BK - 1) / BK;
  const int tiles_mn = m_tiles * n_tiles;
  const long long total_items = static_cast<long long>(tiles_mn) * p.split_k * p.batch;

  // Producer and MMA issuer run warp-converged with the asynchronous instructions themselves under elect.sync:
  // addresses, descriptors and loop state then stay in uniform registers (a lane-0-only branch makes the
  // compiler wrap every UTMALDG / UTCHMMA in a R2UR.BROADCAST loop that costs ~100 cycles per instruction).
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    // CTA2: both CTAs load their own 128 rows of A and their half of B; every byte is credited to the leader's
    // full barrier, which therefore expects the stage of both CTAs
    constexpr int BNL = CTA2 ? BN / 2 : BN;     // B rows staged by this CTA
    auto load = [&](void* dst, const void* tm, uint64_t* bar, int c0, int c1, int c2) {
      if (CTA2) tma_load_3d_2sm(dst, tm, bar, c0, c1, c2); else tma_load_3d(dst, tm, bar, c0, c1, c2);
    };
    for (long long item = worker; item < total_items; item += nworkers) {
      const WorkItem w = decode_item<BN, TM>(item, n_tiles, m_tiles, p.batch, p.split_k, kb_total);
      const int m0 = w.m0 + rank * BM, n0 = w.n0 + rank * BNL, b = w.b, kb0 = w.kb0, kb1 = w.kb1;
      const int ba = p.a_bcast ? 0 : b;
      const int bb = p.b_bcast ? 0 : b;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
        if (rank == 0) mbar_expect_tx(&full_bar[stage], (CTA2 ? 2 : 1) * (Cfg::A_BYTES + Cfg::B_BYTES));
        uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
        uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
        const int k0 = kb * BK;
        if (!A_MN) {
          load(sa, &tmA, &full_bar[stage], k0, m0, ba);
        } else {
#pragma unroll
          for (int i = 0; i < BM / 64; ++i) load(sa + i * (64 * BK * 2), &tmA, &full_bar[stage], m0 + 64 * i, k0, ba);
        }
        if (!B_MN) {
          if (p.b_kchunk) {
            const int chunk = k0 / p.b_kchunk;
            load(sb, &tmB, &full_bar[stage], k0 - chunk * p.b_kchunk, n0, chunk);
          } else {
            load(sb, &tmB, &full_bar[stage], k0, n0, bb);
          }
        } else {
#pragma unroll
          for (int i = 0; i < BNL / 64; ++i) load(sb + i * (64 * BK * 2), &tmB, &full_bar[stage], n0 + 64 * i, k0, bb);
        }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (the leader CTA of a pair)
    constexpr uint32_t idesc = make_idesc_bf16(TM, BN, A_MN, B_MN);
    // per UMMA_K (=16 elements) advance of the descriptor start address, in 16-byte units
    constexpr uint32_t a_adv = A_MN ? (16 * 128) >> 4 : 32 >> 4;
    constexpr uint32_t b_adv = B_MN ? (16 * 128) >> 4 : 32 >> 4;
    constexpr uint32_t lbo = 64 * BK * 2;  // MN-major: next 64-element MN atom
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (long long item = worker; item < total_items; item += nworkers, ++it) {
      const WorkItem w = decode_item<BN, TM>(item, n_tiles, m_tiles, p.batch, p.split_k, kb_total);
      const int kb0 = w.kb0, kb1 = w.kb1;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t adesc =
            make_smem_desc_sw128(smem_u32(smem_a + stage * Cfg::A_BYTES), A_MN ? lbo : 0, 1024);
        const uint64_t bdesc =
            make_smem_desc_sw128(smem_u32(smem_b + stage * Cfg::B_BYTES), B_MN ? lbo : 0, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (CTA2) umma_f16_ss_2sm(d_tmem, adesc + k * a_adv, bdesc + k * b_adv, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_f16_ss(d_tmem, adesc + k * a_adv, bdesc + k * b_adv, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (CTA2) {   // frees the stage / publishes the accumulator in both CTAs
            umma_commit_2sm(&empty_bar[stage]);
            if (kb == kb1 - 1) umma_commit_2sm(&tfull_bar[as]);
          } else {
            umma_commit(&empty_bar[stage]);
            if (kb == kb1 - 1) umma_commit(&tfull_bar[as]);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = ew >> 2;           // which half of the BN columns
    constexpr int HALF_COLS = BN / 2;
    uint32_t aux_uses = 0;
    uint32_t nstaged = 0;   // chunks staged by this warp so far (selects the staging buffer)
    int it = 0;
    for (long long item = worker; item < total_items; item += nworkers, ++it) {
      const WorkItem w = decode_item<BN, TM>(item, n_tiles, m_tiles, p.batch, p.split_k, kb_total);
      const int m0 = w.m0 + rank * BM, n0 = w.n0, b = w.b;   // this CTA's 128 rows of the accumulator, all BN columns
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + half * HALF_COLS;
      const float* bias = p.bias ? p.bias + static_cast<long long>(b) * p.bias_bs : nullptr;
      int sample = 0;
      if (EPI == UMD_EPI_GATE_RES && p.gate && row_ok) sample = sample_of(p.rmap, row);

      if constexpr (STAGED) {
        constexpr int NB = Cfg::OUT_BUFS;
        uint8_t* sbuf0 = smem_out + ew * (NB * 4096);
        const int srow = quad * 32 + lane;
        const int sw = lane & 7;
        const int trow = m0 + quad * 32;   // first row of this warp's 32-row slab
#pragma unroll 1
        for (int c = 0; c < HALF_COLS; c += 64) {
          const int col0 = n0 + half * HALF_COLS + c;
          const bool active = col0 < p.N;  // warp-uniform
          // staging buffer of this chunk; bulk_wait_read<NB-1>: the store that last used it has drained
          uint8_t* sbuf = sbuf0 + (nstaged % NB) * 4096;
          uint8_t* my_row = sbuf + lane * 128;
          if (AUX_TILE && active && lane == 0) {
            bulk_wait_read<NB - 1>();
            mbar_expect_tx(&aux_bar[ew], 4096);
            tma_load_3d(sbuf, &tmAux, &aux_bar[ew], col0, trow, 0);
          }
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(t_row + c, r0);
          tmem_ld_32x32(t_row + c + 32, r1);
          tmem_ld_wait();
          if (active) {
            float v[64];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = __uint_as_float(r0[j]);
              v[32 + j] = __uint_as_float(r1[j]);
            }
            if (!AUX_TILE && bias) {
#pragma unroll
              for (int j = 0; j < 64; j += 4) {
                if (col0 + j < p.N) {
                  const float4 bv = *reinterpret_cast<const float4*>(bias + col0 + j);
                  v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
                }
              }
            }
            if (EPI == UMD_EPI_DGELU) {
              mbar_wait(&aux_bar[ew], (aux_uses++) & 1);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 uu = ld_shared_v4(smem_u32(my_row) + ((j ^ sw) << 4));
                const uint32_t uw[4] = {uu.x, uu.y, uu.z, uu.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  v[8 * j + 2 * q] *= gelu_tanh_grad(bf16_lo(uw[q]));
                  v[8 * j + 2 * q + 1] *= gelu_tanh_grad(bf16_hi(uw[q]));
                }
              }
            } else if (EPI == UMD_EPI_BF16_DELTA) {
              // delta[row, head] = sum_c dO[row, c] O[row, c] over the 64 columns of this chunk (one head: Dh = 64)
              mbar_wait(&aux_bar[ew], (aux_uses++) & 1);
              float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 uu = ld_shared_v4(smem_u32(my_row) + ((j ^ sw) << 4));
                const uint32_t uw[4] = {uu.x, uu.y, uu.z, uu.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  d4[q] = fmaf(v[8 * j + 2 * q], bf16_lo(uw[q]), d4[q]);
                  d4[q] = fmaf(v[8 * j + 2 * q + 1], bf16_hi(uw[q]), d4[q]);
                }
              }
              if (row_ok) reinterpret_cast<float*>(p.out1)[static_cast<long long>(row) * p.ld1 + (col0 >> 6)] = (d4[0] + d4[1]) + (d4[2] + d4[3]);
              __syncwarp();   // every lane has read its operand row before the chunk is overwritten below
            } else {
              if (lane == 0) bulk_wait_read<NB - 1>();
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              st_shared_v4(smem_u32(my_row) + ((j ^ sw) << 4), pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                           pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tmO0, sbuf, col0, trow, b);
              bulk_commit();
            }
            ++nstaged;
            if (EPI == UMD_EPI_GELU) {
              sbuf = sbuf0 + (nstaged % NB) * 4096;
              my_row = sbuf + lane * 128;
              if (lane == 0) bulk_wait_read<NB - 1>();
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float g[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) g[q] = gelu_tanh(v[8 * j + q]);
                st_shared_v4(smem_u32(my_row) + ((j ^ sw) << 4), pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]),
                             pack_bf16x2(g[6], g[7]));
              }
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&tmO1, sbuf, col0, trow, b);
                bulk_commit();
              }
              ++nstaged;
            }
          }
          (void)srow;
        }
      } else {
#pragma unroll 1
      for (int c = 0; c < HALF_COLS; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c, r);
        tmem_ld_wait();
        const int col0 = n0 + half * HALF_COLS + c;
        if (row_ok && col0 < p.N) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const int ncols = min(32, p.N - col0);  // multiple of 8 by contract
        if (EPI != UMD_EPI_ATOMIC && EPI != UMD_EPI_DGELU && bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j < ncols) {
              float4 bv = *reinterpret_cast<const float4*>(bias + col0 + j);
              v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
            }
          }
        }
        if (EPI == UMD_EPI_BF16) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out0) + static_cast<long long>(b) * p.bs0 +
                             static_cast<long long>(row) * p.ld0 + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j < ncols) {
              uint4 pk = make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                                    pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
              *reinterpret_cast<uint4*>(o + j) = pk;
            }
          }
        } else if (EPI == UMD_EPI_F32) {
          float* o = reinterpret_cast<float*>(p.out0) + static_cast<long long>(b) * p.bs0 +
                     static_cast<long long>(row) * p.ld0 + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j < ncols) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        } else if (EPI == UMD_EPI_GELU) {
          __nv_bfloat16* o0 = reinterpret_cast<__nv_bfloat16*>(p.out0) + static_cast<long long>(row) * p.ld0 + col0;
          __nv_bfloat16* o1 = reinterpret_cast<__nv_bfloat16*>(p.out1) + static_cast<long long>(row) * p.ld1 + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j < ncols) {
              uint4 pu = make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                                    pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
              *reinterpret_cast<uint4*>(o0 + j) = pu;
              float g[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) g[q] = gelu_tanh(v[j + q]);
              uint4 pg = make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]),
                                    pack_bf16x2(g[6], g[7]));
              *reinterpret_cast<uint4*>(o1 + j) = pg;
            }
          }
        } else if (EPI == UMD_EPI_GATE_RES) {
          const float* xin = reinterpret_cast<const float*>(p.aux) + static_cast<long long>(row) * p.ldaux + col0;
          float* xo = reinterpret_cast<float*>(p.out1) + static_cast<long long>(row) * p.ld1 + col0;
          const float* gp = p.gate ? p.gate + static_cast<long long>(sample) * p.ldgate + col0 : nullptr;
          __nv_bfloat16* o0 =
              p.out0 ? reinterpret_cast<__nv_bfloat16*>(p.out0) + static_cast<long long>(row) * p.ld0 + col0 : nullptr;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j < ncols) {
              if (o0) {
                uint4 pa = make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                                      pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
                *reinterpret_cast<uint4*>(o0 + j) = pa;
              }
#pragma unroll
              for (int q = 0; q < 8; q += 4) {
                float4 xi = *reinterpret_cast<const float4*>(xin + j + q);
                float4 gv = gp ? *reinterpret_cast<const float4*>(gp + j + q) : make_float4(1.f, 1.f, 1.f, 1.f);
                float4 o;
                o.x = xi.x + gv.x * v[j + q];
                o.y = xi.y + gv.y * v[j + q + 1];
                o.z = xi.z + gv.z * v[j + q + 2];
                o.w = xi.w + gv.w * v[j + q + 3];
                *reinterpret_cast<float4*>(xo + j + q) = o;
              }
            }
          }
        } else if (EPI == UMD_EPI_DGELU) {
          const __nv_bfloat16* up =
              reinterpret_cast<const __nv_bfloat16*>(p.aux) + static_cast<long long>(row) * p.ldaux + col0;
          __nv_bfloat16* o0 = reinterpret_cast<__nv_bfloat16*>(p.out0) + static_cast<long long>(row) * p.ld0 + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j < ncols) {
              uint4 uu = *reinterpret_cast<const uint4*>(up + j);
              uint32_t uw[4] = {uu.x, uu.y, uu.z, uu.w};
              float d[8];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                d[2 * q] = v[j + 2 * q] * gelu_tanh_grad(bf16_lo(uw[q]));
                d[2 * q + 1] = v[j + 2 * q + 1] * gelu_tanh_grad(bf16_hi(uw[q]));
              }
              uint4 pd = make_uint4(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]),
                                    pack_bf16x2(d[6], d[7]));
              *reinterpret_cast<uint4*>(o0 + j) = pd;
            }
          }
        } else if (EPI == UMD_EPI_ATOMIC) {
          float* o = reinterpret_cast<float*>(p.out0) + static_cast<long long>(b) * p.bs0 +
                     static_cast<long long>(row) * p.ld0 + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j < ncols) red_add_v4(o + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
        }  // row_ok && col0 < N
        __syncwarp();
      }
      }  // !STAGED
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_leader(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]);
      }
    }
    if (STAGED && lane == 0) bulk_wait<0>();  // the staging buffer must outlive the last TMA store's read
    (void)aux_uses;
    (void)nstaged;
  }

  __syncwarp();
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();   // neither CTA of a pair leaves while its peer may still read it
  if (warp == 2) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------
extern long long g_launch_count;

template <int BN, bool A_MN, bool B_MN, int EPI, bool CTA2>
static int launch_gemm_cfg(const CUtensorMap* tm, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CTA2>;
  auto kern = gemm_kernel<BN, A_MN, B_MN, EPI, CTA2>;
  static bool configured = false;
  if (!configured) {
    UMD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tile_m = CTA2 ? 2 * BM : BM;
  const int m_tiles = ceil_div(p.M, tile_m), n_tiles = ceil_div(p.N, BN);
  const long long items = static_cast<long long>(m_tiles) * n_tiles * p.split_k * p.batch;
  const int workers_max = CTA2 ? sm_count() / 2 : sm_count();
  const int workers = static_cast<int>(items < workers_max ? items : workers_max);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(CTA2 ? 2 * workers : workers);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTA2 ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  UMD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tm[0], tm[1], tm[2], tm[3], tm[4], p));
  ++g_launch_count;
  return UMD_OK;
}

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm_t(const CUtensorMap* tm, const GemmParams& p, cudaStream_t stream) {
  if (BN == 256 && p.cta2) return launch_gemm_cfg<256, A_MN, B_MN, EPI, true>(tm, p, stream);
  return launch_gemm_cfg<BN, A_MN, B_MN, EPI, false>(tm, p, stream);
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm_epi(int epi, const CUtensorMap* tm, const GemmParams& p, cudaStream_t s) {
  switch (epi) {
    case UMD_EPI_BF16: return launch_gemm_t<BN, A_MN, B_MN, UMD_EPI_BF16>(tm, p, s);
    case UMD_EPI_F32: return launch_gemm_t<BN, A_MN, B_MN, UMD_EPI_F32>(tm, p, s);
    case UMD_EPI_ATOMIC: return launch_gemm_t<BN, A_MN, B_MN, UMD_EPI_ATOMIC>(tm, p, s);
    default: break;
  }
  if (!A_MN) {
    // fused activation tails only exist for activations-as-A (forward / dgrad) GEMMs
    switch (epi) {
      case UMD_EPI_GELU: return launch_gemm_t<BN, false, B_MN, UMD_EPI_GELU>(tm, p, s);
      case UMD_EPI_GATE_RES: return launch_gemm_t<BN, false, B_MN, UMD_EPI_GATE_RES>(tm, p, s);
      case UMD_EPI_DGELU: return launch_gemm_t<BN, false, B_MN, UMD_EPI_DGELU>(tm, p, s);
      case UMD_EPI_BF16_DELTA:
        if (BN >= 128) return launch_gemm_t<(BN >= 128 ? BN : 128), false, B_MN, UMD_EPI_BF16_DELTA>(tm, p, s);
        break;
      default: break;
    }
  }
  set_error("umd_gemm_bf16: unsupported epilogue %d for this operand layout", epi);
  return UMD_ERR_UNSUPPORTED;
}

template <int BN>
static int launch_gemm_bn(bool a_mn, bool b_mn, int epi, const CUtensorMap* tm, const GemmParams& p, cudaStream_t s) {
  if (!a_mn && b_mn) return launch_gemm_epi<BN, false, true>(epi, tm, p, s);
  if (!a_mn && !b_mn) return launch_gemm_epi<BN, false, false>(epi, tm, p, s);
  if (a_mn && b_mn) return launch_gemm_epi<BN, true, true>(epi, tm, p, s);
  set_error("umd_gemm_bf16: A MN-major with B K-major is not instantiated");
  return UMD_ERR_UNSUPPORTED;
}

int gemm_bf16(const umd_gemm_args& a, cudaStream_t stream) {
  UMD_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0 && a.batch > 0, "umd_gemm_bf16: empty problem M=%d N=%d K=%d batch=%d",
              a.M, a.N, a.K, a.batch);
  UMD_REQUIRE(a.N % 8 == 0, "umd_gemm_bf16: N=%d must be a multiple of 8", a.N);
  UMD_REQUIRE(a.lda % 8 == 0 && a.ldb % 8 == 0, "umd_gemm_bf16: lda/ldb must be multiples of 8 elements");
  UMD_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0,
              "umd_gemm_bf16: operands must be 16-byte aligned");
  const int split_k = a.split_k > 1 ? a.split_k : 1;
  UMD_REQUIRE(split_k == 1 || a.epi == UMD_EPI_ATOMIC, "umd_gemm_bf16: split_k needs UMD_EPI_ATOMIC");
  const int kb_total = ceil_div(a.K, BK);
  GemmParams p;
  p.M = a.M; p.N = a.N; p.K = a.K; p.batch = a.batch;
  p.split_k = split_k < kb_total ? split_k : kb_total;
  p.a_bcast = (a.a_bs == 0); p.b_bcast = (a.b_bs == 0);
  p.b_kchunk = a.b_kchunk;
  if (a.b_kchunk) {
    UMD_REQUIRE(!a.b_mn && a.batch == 1 && a.b_kchunk % BK == 0 && a.K % a.b_kchunk == 0 && a.b_bs > 0,
                "umd_gemm_bf16: b_kchunk needs K-major B, batch 1, chunk %% 64 == 0, K %% chunk == 0");
    p.b_bcast = 0;
  }
  p.out0 = a.out0; p.ld0 = a.ld0; p.bs0 = a.bs0;
  p.out1 = a.out1; p.ld1 = a.ld1;
  p.bias = a.bias; p.bias_bs = a.bias_bs;
  p.aux = a.aux; p.ldaux = a.ldaux;
  p.gate = a.gate; p.ldgate = a.ldgate;
  p.rmap.split_row = a.split_row; p.rmap.s0 = a.s0 > 0 ? a.s0 : 1; p.rmap.s1 = a.s1 > 0 ? a.s1 : 1; p.rmap.n0 = a.n0;

  int bn = 256;
  if (a.N <= 64) bn = 64;
  else if (a.N <= 128) bn = 128;
  else if (a.N % 256 != 0 && a.N % 128 == 0 && a.N < 1024) bn = 128;

  {
    static int use2 = -1;
    if (use2 < 0) {
      const char* e = getenv("UMD_GEMM_CTA2");
      use2 = e ? atoi(e) : 1;
    }
    p.cta2 = (use2 && bn == 256 && a.M >= 256) ? 1 : 0;
  }
  const int b_box = p.cta2 ? bn / 2 : bn;   // rows of B staged per CTA
  CUtensorMap tm[5];
  CUtensorMap &tmA = tm[0], &tmB = tm[1];
  const uint64_t a_batch = p.a_bcast ? 1 : a.batch, b_batch = p.b_bcast ? 1 : a.batch;
  if (!a.a_mn) UMD_TRY(make_tmap_bf16(&tmA, a.A, a.K, a.M, a_batch, a.lda, a.a_bs, BM));
  else         UMD_TRY(make_tmap_bf16(&tmA, a.A, a.M, a.K, a_batch, a.lda, a.a_bs, BK));
  if (a.b_kchunk) UMD_TRY(make_tmap_bf16(&tmB, a.B, a.b_kchunk, a.N, a.K / a.b_kchunk, a.ldb, a.b_bs, b_box));
  else if (!a.b_mn) UMD_TRY(make_tmap_bf16(&tmB, a.B, a.K, a.N, b_batch, a.ldb, a.b_bs, b_box));
  else         UMD_TRY(make_tmap_bf16(&tmB, a.B, a.N, a.K, b_batch, a.ldb, a.b_bs, BK));

  tm[2] = tm[3] = tm[4] = tmA;  // placeholders for the epilogues that do not use them
  UMD_REQUIRE(a.epi != UMD_EPI_BF16_DELTA || (bn >= 128 && a.N % 64 == 0 && a.out1 && a.ld1 > 0 && a.batch == 1),
              "umd_gemm_bf16: the delta epilogue needs N %% 64 == 0 (N > 64), out1 with ld1 = groups per row, batch 1");
  if (bn >= 128 && (a.epi == UMD_EPI_BF16 || a.epi == UMD_EPI_GELU || a.epi == UMD_EPI_DGELU || a.epi == UMD_EPI_BF16_DELTA)) {
    UMD_REQUIRE(a.out0 && (reinterpret_cast<uintptr_t>(a.out0) & 15) == 0 && a.ld0 % 8 == 0 && a.bs0 % 8 == 0,
                "umd_gemm_bf16: bf16 outputs must be 16-byte aligned with ld0 / bs0 multiples of 8");
    UMD_TRY(make_tmap_bf16(&tm[2], a.out0, a.N, a.M, a.batch, a.ld0, a.bs0, 32));
    if (a.epi == UMD_EPI_GELU) {
      UMD_REQUIRE(a.out1 && (reinterpret_cast<uintptr_t>(a.out1) & 15) == 0 && a.ld1 % 8 == 0 && a.batch == 1,
                  "umd_gemm_bf16: GELU epilogue needs an aligned out1 with ld1 %% 8 == 0 and batch 1");
      UMD_TRY(make_tmap_bf16(&tm[3], a.out1, a.N, a.M, 1, a.ld1, 0, 32));
    }
    if (a.epi == UMD_EPI_DGELU || a.epi == UMD_EPI_BF16_DELTA) {
      UMD_REQUIRE(a.aux && (reinterpret_cast<uintptr_t>(a.aux) & 15) == 0 && a.ldaux % 8 == 0 && a.batch == 1,
                  "umd_gemm_bf16: DGELU / delta epilogues need an aligned aux with ldaux %% 8 == 0 and batch 1");
      UMD_TRY(make_tmap_bf16(&tm[4], a.aux, a.N, a.M, 1, a.ldaux, 0, 32));
    }
  }
  {
    // measurement aid (tools/traffic_table.py): UMD_GEMM_LOG=<file> appends one line per launch, in launch order, so that an
    // ncu launch list of the same process can be joined with the problem shapes
    static FILE* logf = nullptr;
    static int log_state = 0;
    if (log_state == 0) {
      const char* e = getenv("UMD_GEMM_LOG");
      logf = (e && *e) ? fopen(e, "a") : nullptr;
      log_state = 1;
    }
    if (logf) {
      fprintf(logf, "%d %d %d %d %d %d %d %d %d %d %d %d\n", a.M, a.N, a.K, a.batch, a.a_mn, a.b_mn, a.epi, p.split_k, bn, p.cta2,
              p.a_bcast, p.b_bcast);
      fflush(logf);
    }
  }
  ProfScope prof(a.a_mn ? PC_GEMM_WGRAD : PC_GEMM, 2.0 * a.M * static_cast<double>(a.N) * a.K * a.batch, stream);
  switch (bn) {
    case 256: return launch_gemm_bn<256>(a.a_mn, a.b_mn, a.epi, tm, p, stream);
    case 128: return launch_gemm_bn<128>(a.a_mn, a.b_mn, a.epi, tm, p, stream);
    default: return launch_gemm_bn<64>(a.a_mn, a.b_mn, a.epi, tm, p, stream);
  }
}

}  // namespace umd

extern "C" int umd_gemm_bf16(const umd_gemm_args* args, umd_stream_t stream) {
  if (!args) {
    umd::set_error("umd_gemm_bf16: null args");
    return UMD_ERR_INVALID;
  }
  return umd::gemm_bf16(*args, static_cast<cudaStream_t>(stream));
}
