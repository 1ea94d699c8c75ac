// Fused multi-head self-attention on tcgen05 / TMEM for sm_100a, forward and backward.
//
// Semantics: flax MultiHeadDotProductAttention as called at big_vision/models/vit.py:82-87 —
// softmax((q / sqrt(Dh)) k^T) v per (sample, head), no mask, no dropout; backward per SURVEY.md App. E step 7.
// Every recipe of the reference has at most 260 tokens per sample, so one CTA owns one
// (sample, head): Q, K, V (and dO) of that head are fetched once by TMA into 128B-swizzled shared memory and
// the S x S score matrix never leaves the SM.
//
//   forward : per 128-query tile  S = Q K^T -> TMEM;  softmax warps (thread = TMEM lane = query row) read S,
//             write P (bf16) to shared memory in the canonical K-major swizzled layout;  O = P V -> TMEM;
//             epilogue scales by 1 / rowsum and stores O (bf16) and the log-sum-exp.
//   backward: per 128-key tile j and 128-query tile i (in 64-query halves):  S^T = K_j Q_i^T and
//             dP^T = V_j dO_i^T -> TMEM;  the softmax warps (thread = key row) form P^T and
//             dS^T = P^T (dP^T - delta) and write both (bf16) to shared memory;  dV_j += P^T dO_i,
//             dK_j += dS^T Q_i (K-major A) and dQ_i += dS K_j (the same tile read as an MN-major A) accumulate
//             in TMEM; dK/dV leave after the query loop, dQ after the key loop.
//
// Warp roles: warp 0 TMA producer (one lane), warp 1 MMA issuer, warp 2 TMEM allocator, warps 4.. softmax / epilogue:
// warp w owns TMEM lanes 32*(w%4) .. +31 and one part of the score columns.  Softmax is bound by MUFU.EX2 (16 lanes /
// clk / SM) and by the TMEM read rate (~100-130 B / clk / SM), and its dependent chain load -> FFMA -> EX2 -> pack ->
// store only reaches the MUFU rate with four warps per SM sub-partition (tools/ubench/pipes.cu), hence 16 softmax
// warps per CTA (or two 8-warp CTAs per SM for short sequences).  Warps 0, 2, 3 are idle after set-up and take the
// one-row tail of 257-token sequences (see split_tail).  Measurements and dead ends: profiles/r01_notes.md.
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace umd {

extern long long g_launch_count;

long long* g_attn_timeline = nullptr;

namespace {

constexpr int AT_DH = 64;
constexpr int AT_MAX_S = 272;
constexpr int AT_ROW = 128;            // bytes per 64-element bf16 row
constexpr int AT_SLAB = 128 * AT_ROW;  // [128 rows][64 cols] bf16 = 16 KB
constexpr int AT_BWD_THREADS = 640;     // backward: 4 control warps + 16 softmax warps (four per TMEM lane quadrant)
constexpr int AT_BWD_SM = 512;          // backward: softmax / epilogue threads
constexpr int AT_STAT_N = 384;         // per-query statistics padded to three 128-query tiles
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#define TL(slot, idx) do { if (tl_on) p.tl[(slot) * 64 + (idx)] = clock64(); } while (0)

__device__ __forceinline__ void bar_softmax() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
template <int NT>
__device__ __forceinline__ void bar_softmax_n() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// rows [0, rows) of one sample (tensor map dims: column, row in sample, sample; rows >= S are zero-filled).
// tm16 is the tensor map whose box holds the rows % 128 tail rows (a multiple of 16).
__device__ __forceinline__ void load_rows(uint8_t* dst, const CUtensorMap* tm128, const CUtensorMap* tm16, uint64_t* bar,
                                          int col, int sample, int rows) {
  int r = 0;
#pragma unroll 1
  for (; r + 128 <= rows; r += 128) tma_load_3d(dst + r * AT_ROW, tm128, bar, col, r, sample);
  if (r < rows) tma_load_3d(dst + r * AT_ROW, tm16, bar, col, r, sample);  // one box of rows % 128 rows
}

// rows [r0, r1) of one sample (r0 a multiple of 128), see load_rows
__device__ __forceinline__ void load_rows_from(uint8_t* dst, const CUtensorMap* tm128, const CUtensorMap* tm16, uint64_t* bar,
                                               int col, int sample, int r0, int r1) {
  int r = r0;
#pragma unroll 1
  for (; r + 128 <= r1; r += 128) tma_load_3d(dst + r * AT_ROW, tm128, bar, col, r, sample);
  if (r < r1) tma_load_3d(dst + r * AT_ROW, tm16, bar, col, r, sample);
}

// L2 prefetch of rows [0, rows) of one sample: issued for the (sample, head) that the CTA one resident wave ahead
// will load, so that its TMA loads hit L2 instead of paying the HBM round trip at the head of a serial chain
__device__ __forceinline__ void prefetch_rows(const CUtensorMap* tm128, const CUtensorMap* tm16, int col, int sample, int rows) {
  int r = 0;
#pragma unroll 1
  for (; r + 128 <= rows; r += 128) tma_prefetch_l2_3d(tm128, col, r, sample);
  if (r < rows) tma_prefetch_l2_3d(tm16, col, r, sample);
}

struct FwdParams {
  __nv_bfloat16* out;
  float* lse;
  int row_base;  // first row of the segment
  int S, SP, H;
  int nqt;       // 128-query tiles that run on the tensor cores
  int ntail;     // trailing query rows (S = 128 nqt + ntail, ntail <= AT_TAIL) computed on the idle control warps
  int o_col, tmem_cols;
  int ahead;     // L2 prefetch of this CTA's next item on / off
  int nitems;    // (sample, head) pairs of the segment
  float scale_log2;
  int tl_cta;
  long long* tl;  // optional timeline buffer (tools/attn_timeline.py)
};

constexpr int AT_TAIL = 4;        // largest query / key tail handled outside the 128-row tiles
constexpr int AT_TAIL_THREADS = 96;

// W score columns [c, c + W) of one query row: running maximum
template <int W, bool MASK>
__device__ __forceinline__ void fwd_max(const uint32_t* v, int c, int S, float (&m)[4]) {
#pragma unroll
  for (int j = 0; j < W; ++j) {
    float x = __uint_as_float(v[j]);
    if (MASK) x = (c + j < S) ? x : -INFINITY;
    m[j & 3] = fmaxf(m[j & 3], x);
  }
}
// ... p = exp2(s * scale*log2e - max) -> partial row sums and the bf16 P tile (K-major, 128B swizzle)
template <int W, bool MASK>
__device__ __forceinline__ void fwd_exp(const uint32_t* v, int c, int S, float sl2, float ms, float (&sum)[4], uint32_t sp_row,
                                        int r7) {
#pragma unroll
  for (int j = 0; j < W; j += 8) {
    float e[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float x = ex2(fmaf(__uint_as_float(v[j + q]), sl2, -ms));
      if (MASK) x = (c + j + q < S) ? x : 0.f;
      e[q] = x;
      sum[q & 3] += x;
    }
    const int cc = c + j;
    st_shared_v4(sp_row + ((cc >> 6) << 14) + ((((cc >> 3) & 7) ^ r7) << 4), pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]),
                 pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
  }
}

// Softmax needs MUFU.EX2 (16 lanes / clk / SM) behind a dependent chain TMEM load -> FFMA -> EX2 -> pack -> store; it
// only reaches the MUFU rate with four warps per SM sub-partition (tools/ubench/pipes.cu: 64 cycles per 8 EX2 at four
// warps, 93 at two).  NQ = softmax warps per TMEM lane quadrant; each owns 1/NQ of the score columns of its 32 rows.
//   TWO  (NQ = 2, 384 threads): the CTA needs at most 256 TMEM columns and half of the shared memory, two CTAs share
//        an SM (register budget 85 per thread) and supply the four warps per sub-partition between them;
//   !TWO (NQ = 4, 640 threads): one CTA per SM.
template <bool TWO>
__global__ void __launch_bounds__(TWO ? 384 : 640, TWO ? 2 : 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm16,
                   const __grid_constant__ CUtensorMap tmo, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // Persistent: the CTA walks over (sample, head) items blockIdx.x, blockIdx.x + gridDim.x, ...  TMEM, barriers and
  // roles are set up once; barrier phases run on (k = item count of this CTA, g = k * nqt + tile).  A CTA per item
  // paid ~2 us of launch, set-up and tear-down for every ~8 us of work.
  const int S = p.S, SP = p.SP, nqt = p.nqt;
  const int D = p.H * AT_DH;
  const int nslab = (SP + 63) >> 6;
  const int qrows = max(nqt * 128, SP);
  uint8_t* sQ = smem;                      // qrows rows (rows >= SP stay unwritten: they only feed unused lanes)
  uint8_t* sK = sQ + qrows * AT_ROW;       // SP rows
  uint8_t* sV = sK + SP * AT_ROW;          // SP rows
  uint8_t* sP = sV + SP * AT_ROW;          // nslab slabs [128 q][64 kv]
  constexpr int NQ = TWO ? 2 : 4;
  constexpr int SMT = 128 * NQ;                                    // softmax / epilogue threads
  uint8_t* sStage = sP + nslab * AT_SLAB;  // [128 q][64] output tile on its way to the TMA store
  float* sMax = reinterpret_cast<float*>(sStage + AT_SLAB);        // [NQ parts][128 rows]
  float* sSum = sMax + 512;                                        // [2 tiles in flight][NQ parts][128 rows]
  float* sTs = sSum + 1024;                                        // [AT_MAX_S] tail-row scores / probabilities
  float* sTr = sTs + AT_MAX_S;                                     // [3][64] tail-row partial outputs, [8] reductions
  uint64_t* bars = reinterpret_cast<uint64_t*>(sTr + 3 * 64 + 16);
  uint64_t* bar_k = bars + 0;    // K and the first query tile have landed
  uint64_t* bar_v = bars + 1;    // V and the remaining query rows have landed
  uint64_t* bar_s = bars + 2;    // scores of the current tile are in TMEM
  uint64_t* bar_p = bars + 3;    // P of the current tile is in shared memory (and S has been read)
  uint64_t* bar_o = bars + 4;    // [2] P V of tile g is in TMEM accumulator g & 1
  uint64_t* bar_free = bars + 6; // every MMA of the item has completed: Q, K, V may be overwritten (once per item)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // loads of one item: two completion groups so that Q_0 K^T can start once half of the bytes are in; then an L2
  // prefetch of the item this CTA takes next
  auto issue_loads = [&](int item) {
    const int h = item % p.H, sample = item / p.H;
    const int q0rows = min(128, SP);
    mbar_expect_tx(bar_k, static_cast<uint32_t>(SP + q0rows) * AT_ROW);
    load_rows_from(sK, &tm128, &tm16, bar_k, D + h * AT_DH, sample, 0, SP);
    load_rows_from(sQ, &tm128, &tm16, bar_k, h * AT_DH, sample, 0, q0rows);
    mbar_expect_tx(bar_v, static_cast<uint32_t>(2 * SP - q0rows) * AT_ROW);
    load_rows_from(sV, &tm128, &tm16, bar_v, 2 * D + h * AT_DH, sample, 0, SP);
    if (SP > q0rows) load_rows_from(sQ, &tm128, &tm16, bar_v, h * AT_DH, sample, 128, SP);
    const int nxt = item + static_cast<int>(gridDim.x);
    if (p.ahead > 0 && nxt < p.nitems) {
      const int h2 = nxt % p.H, s2 = nxt / p.H;
      prefetch_rows(&tm128, &tm16, D + h2 * AT_DH, s2, SP);
      prefetch_rows(&tm128, &tm16, h2 * AT_DH, s2, SP);
      prefetch_rows(&tm128, &tm16, 2 * D + h2 * AT_DH, s2, SP);
    }
  };
  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer: the first item's loads go out
    // BEFORE the set-up barrier (this thread owns the two load barriers), so they are in flight while TMEM is
    // allocated and the CTA synchronises
    tma_prefetch_desc(&tm128);
    tma_prefetch_desc(&tm16);
    tma_prefetch_desc(&tmo);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    fence_barrier_init();
    issue_loads(blockIdx.x);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar_s, 1);
    mbar_init(&bar_o[0], 1);
    mbar_init(&bar_o[1], 1);
    mbar_init(bar_free, 1);
    mbar_init(bar_p, SMT);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-converged; the tcgen05
    // instructions themselves run under elect.sync so that descriptors stay in uniform registers)
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, false, true);
    const uint64_t v_desc = make_smem_desc_sw128(smem_u32(sV), 8192, 1024);
    const uint64_t p_desc = make_smem_desc_sw128(smem_u32(sP), 0, 1024);
    const int ksteps = SP >> 4;
    bool tl_on = false;
    // S = Q_i K^T for query tile i (two column chunks when SP > 256).  (Descriptors of the two score chunks are
    // rebuilt per tile: a hoisted form of this short sequence produced wrong columns >= 256 on hardware.)
    auto issue_qk = [&](int i) {
      const uint64_t adesc = make_smem_desc_sw128(smem_u32(sQ + i * 128 * AT_ROW), 0, 1024);
      if (elect_one()) {
        for (int n0 = 0; n0 < SP; n0 += 256) {
          const int n = min(256, SP - n0);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sK + n0 * AT_ROW), 0, 1024);
          const uint32_t idesc = make_idesc_bf16(128, n, false, false);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + n0, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(bar_s);
      }
      __syncwarp();
    };
    int k = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++k) {
    tl_on = p.tl != nullptr && item == p.tl_cta && lane == 0;
    // (the S columns are free: this warp waited for bar_p of the previous item's last tile before its P V)
    mbar_wait(bar_k, k & 1);
    tc_fence_after();
    TL(0, 0);
    TL(1, 0);
    issue_qk(0);
    TL(2, 0);
    mbar_wait(bar_v, k & 1);
    tc_fence_after();
    for (int i = 0; i < nqt; ++i) {
      const int g = k * nqt + i;
      // bar_p(i): every softmax thread has finished reading S(i) and writing P(i).  The next tile's scores go first
      // (the softmax warps start on them while P V runs); O_i goes to accumulator i & 1, whose previous content
      // (O_{i-2}) was read by the epilogue that precedes the arrival on bar_p(i-1) in program order.
      mbar_wait(bar_p, g & 1);
      tc_fence_after();
      TL(3, i);
      if (i + 1 < nqt) {
        TL(1, i + 1);
        issue_qk(i + 1);
        TL(2, i + 1);
      }
      const uint32_t o_acc = tmem_base + p.o_col + 64 * (g & 1);
      if (elect_one()) {
        uint64_t pa = p_desc, vb = v_desc;
        int kk = 0;
#pragma unroll 1
        for (; kk + 4 <= ksteps; kk += 4) {  // one 64-key slab of P per iteration
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(o_acc, pa + 2 * k, vb + static_cast<uint64_t>(k * (16 * AT_ROW / 16)), idesc_pv, (kk + k) > 0 ? 1u : 0u);
          pa += AT_SLAB / 16;
          vb += 4 * (16 * AT_ROW / 16);
        }
        for (int k = 0; kk < ksteps; ++kk, ++k)
          umma_f16_ss(o_acc, pa + 2 * k, vb + static_cast<uint64_t>(k * (16 * AT_ROW / 16)), idesc_pv, kk > 0 ? 1u : 0u);
        umma_commit(&bar_o[g & 1]);
        if (i == nqt - 1) umma_commit(bar_free);
      }
      __syncwarp();
    }
    }  // items
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + epilogue
    const int quad = warp & 3;
    const int hf = (warp - 4) >> 2;  // which part of the score columns (and of the 64 output columns)
    const int r = quad * 32 + lane;  // row inside the query tile = TMEM lane
    const int r7 = r & 7;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t sp_row = smem_u32(sP) + r * AT_ROW;
    const float sl2 = p.scale_log2;
    // the SP / 16 column groups are dealt out as evenly as possible
    const int units = SP >> 4, ubase = units / NQ, urem = units % NQ;
    const int cb = 16 * (hf * ubase + min(hf, urem)), ce = cb + 16 * (ubase + (hf < urem ? 1 : 0));
    const int tid = threadIdx.x - 128;
    bool tl_on = false;
    const uint32_t stage_row = smem_u32(sStage) + r * AT_ROW;
    int h = 0, sample = 0, row0 = 0, k = 0;
    // epilogue of tile i (global tile number g): O / rowsum -> bf16 -> staging tile -> TMA store (clipped at S), lse
    auto epilogue = [&](int i, int g, float ms) {
      constexpr int OC = 64 / NQ;          // output columns per part
      if (tid == 0) bulk_wait_read<0>();   // the previous store has finished reading the staging tile
      mbar_wait(&bar_o[g & 1], (g >> 1) & 1);
      tc_fence_after();
      TL(8, i);
      uint32_t o0[OC];
      if constexpr (OC == 32) {
        tmem_ld_32x32(t_lane + p.o_col + 64 * (g & 1) + OC * hf, o0);
        tmem_ld_wait_dep32(o0);
      } else {
        tmem_ld_32x16(t_lane + p.o_col + 64 * (g & 1) + OC * hf, o0);
        tmem_ld_wait_dep16(o0);
      }
      tc_fence_before();
      bar_softmax_n<SMT>();                // partial sums of all parts are visible; staging tile is free
      float sum = 0.f;
#pragma unroll
      for (int q = 0; q < NQ; ++q) sum += sSum[(g & 1) * 512 + q * 128 + r];
      const int row = i * 128 + r;
      const float inv = 1.f / sum;
#pragma unroll
      for (int j = 0; j < OC; j += 8) {
        st_shared_v4(stage_row + ((((OC / 8) * hf + (j >> 3)) ^ r7) << 4),
                     pack_bf16x2(__uint_as_float(o0[j]) * inv, __uint_as_float(o0[j + 1]) * inv),
                     pack_bf16x2(__uint_as_float(o0[j + 2]) * inv, __uint_as_float(o0[j + 3]) * inv),
                     pack_bf16x2(__uint_as_float(o0[j + 4]) * inv, __uint_as_float(o0[j + 5]) * inv),
                     pack_bf16x2(__uint_as_float(o0[j + 6]) * inv, __uint_as_float(o0[j + 7]) * inv));
      }
      fence_proxy_async();
      bar_softmax_n<SMT>();
      if (tid == 0) {
        tma_store_3d(&tmo, sStage, h * AT_DH, i * 128, sample);
        bulk_commit();
      }
      if (row < S && p.lse && hf == 0) p.lse[static_cast<long long>(row0 + row) * p.H + h] = (ms + log2f(sum)) * LN2;
      TL(9, i);
    };
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++k) {
    h = item % p.H;
    sample = item / p.H;
    row0 = p.row_base + sample * S;
    tl_on = p.tl != nullptr && item == p.tl_cta && tid == 0;
    TL(4, 0);
    float ms_prev = 0.f;
    for (int i = 0; i < nqt; ++i) {
      const int g = k * nqt + i;
      mbar_wait(bar_s, g & 1);
      tc_fence_after();
      TL(5, i);
      // pass 1: maximum of the raw logits over this thread's columns (32 at a time; a 16-column remainder last)
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
      for (int c = cb; c < ce; c += 32) {
        uint32_t v[32];
        if (c + 32 <= ce) {
          tmem_ld_32x32(t_lane + c, v);
          tmem_ld_wait_dep32(v);
          if (c + 32 <= S) fwd_max<32, false>(v, c, S, m4); else fwd_max<32, true>(v, c, S, m4);
        } else {
          tmem_ld_32x16(t_lane + c, reinterpret_cast<uint32_t(&)[16]>(v));
          tmem_ld_wait_dep16(reinterpret_cast<uint32_t(&)[16]>(v));
          if (c + 16 <= S) fwd_max<16, false>(v, c, S, m4); else fwd_max<16, true>(v, c, S, m4);
        }
      }
      float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      sMax[hf * 128 + r] = m;
      TL(6, i);
      bar_softmax_n<SMT>();
#pragma unroll
      for (int q = 1; q < NQ; ++q) m = fmaxf(m, sMax[((hf + q) % NQ) * 128 + r]);
      const float ms = m * sl2;
      // P V of the previous tile reads the P slabs that pass 2 is about to overwrite (its Q K^T successor was issued
      // ahead of it, so S(i) can be ready before P V(i-1) has finished): wait for its completion first
      if (i > 0) mbar_wait(&bar_o[(g - 1) & 1], ((g - 1) >> 1) & 1);
      // pass 2: p = exp2(s * scale*log2e - max), partial row sum, bf16 P -> shared memory
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = cb; c < ce; c += 32) {
        uint32_t v[32];
        if (c + 32 <= ce) {
          tmem_ld_32x32(t_lane + c, v);
          tmem_ld_wait_dep32(v);
          if (c + 32 <= S) fwd_exp<32, false>(v, c, S, sl2, ms, s4, sp_row, r7); else fwd_exp<32, true>(v, c, S, sl2, ms, s4, sp_row, r7);
        } else {
          tmem_ld_32x16(t_lane + c, reinterpret_cast<uint32_t(&)[16]>(v));
          tmem_ld_wait_dep16(reinterpret_cast<uint32_t(&)[16]>(v));
          if (c + 16 <= S) fwd_exp<16, false>(v, c, S, sl2, ms, s4, sp_row, r7); else fwd_exp<16, true>(v, c, S, sl2, ms, s4, sp_row, r7);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar_p);
      sSum[(g & 1) * 512 + hf * 128 + r] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      TL(7, i);
      // the epilogue of the PREVIOUS tile runs here, behind this tile's softmax: its P V finished long ago, and this
      // tile's P V (and the next tile's Q K^T) proceed meanwhile
      if (i > 0) epilogue(i - 1, g - 1, ms_prev);
      ms_prev = ms;
    }
    // last tile of the item: straight away (the next item's loads are only just being issued)
    epilogue(nqt - 1, k * nqt + nqt - 1, ms_prev);
    }  // items
    if (tid == 0) bulk_wait_read<0>();
  } else {
    // ------------------------------------------------------------------ warps 0, 2, 3: the producer thread issues the
    // loads of the following items; with a query tail (one-CTA-per-SM variant only: the launcher sends every shape with
    // a tail there) all 96 threads compute the tail rows on CUDA cores.
    // S = 128 nqt + ntail with ntail <= AT_TAIL (every decoder of the reference has 257 tokens, the label-conditioned
    // encoder 260): a third query tile would run the whole softmax for one to four live rows.  The idle control
    // warps compute those rows from the K / V / Q tiles already in shared memory instead.
    const int tt = (warp == 0 ? 0 : warp - 1) * 32 + lane;   // 0..95
    const int tw = tt >> 5;
    const float sl2 = p.scale_log2;
    const uint32_t q_u = smem_u32(sQ), k_u = smem_u32(sK), v_u = smem_u32(sV);
    const int ntail = TWO ? 0 : p.ntail;
    int k = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++k) {
    if (warp == 0 && lane == 0 && k > 0) {
      // Q, K, V of the previous item are dead once the P V of its last tile has completed (the tail threads passed
      // their last named barrier of that item before this point).  A barrier of its own, completed once per item:
      // waiting on bar_o's parity from here could alias with an earlier phase.
      mbar_wait(bar_free, (k - 1) & 1);
      issue_loads(item);
    }
    __syncwarp();
    if (ntail == 0) continue;
    const int h = item % p.H, sample = item / p.H;
    const int row0 = p.row_base + sample * S;
    mbar_wait(bar_k, k & 1);
    mbar_wait(bar_v, k & 1);
    for (int t = 0; t < p.ntail; ++t) {
      const int qrow = nqt * 128 + t;
      // scores of query qrow against every key: one key per thread, 64-long dot product on packed bf16 pairs
      uint4 qv[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) qv[c] = ld_shared_v4(q_u + qrow * AT_ROW + ((c ^ (qrow & 7)) << 4));
      float mx = -INFINITY;
      for (int k = tt; k < SP; k += AT_TAIL_THREADS) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 kv = ld_shared_v4(k_u + k * AT_ROW + ((c ^ (k & 7)) << 4));
          acc += bf16_lo(qv[c].x) * bf16_lo(kv.x) + bf16_hi(qv[c].x) * bf16_hi(kv.x) + bf16_lo(qv[c].y) * bf16_lo(kv.y) +
                 bf16_hi(qv[c].y) * bf16_hi(kv.y) + bf16_lo(qv[c].z) * bf16_lo(kv.z) + bf16_hi(qv[c].z) * bf16_hi(kv.z) +
                 bf16_lo(qv[c].w) * bf16_lo(kv.w) + bf16_hi(qv[c].w) * bf16_hi(kv.w);
        }
        const float sc = (k < S) ? acc * sl2 : -INFINITY;
        sTs[k] = sc;
        mx = fmaxf(mx, sc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) sTr[192 + tw] = mx;
      bar_sync_named(2, AT_TAIL_THREADS);
      mx = fmaxf(fmaxf(sTr[192], sTr[193]), sTr[194]);
      float sum = 0.f;
      for (int k = tt; k < SP; k += AT_TAIL_THREADS) {
        const float e = ex2(sTs[k] - mx);   // exp2(-inf) = 0 for the padded keys
        sTs[k] = e;
        sum += e;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) sTr[196 + tw] = sum;
      bar_sync_named(2, AT_TAIL_THREADS);
      sum = sTr[196] + sTr[197] + sTr[198];
      // O[qrow] = P V: lane owns output columns 2 lane, 2 lane + 1; warp tw takes keys tw, tw + 3, ...
      float o0 = 0.f, o1 = 0.f;
      for (int k = tw; k < S; k += 3) {
        const float pk = sTs[k];
        const uint32_t vv = ld_shared_b32(v_u + k * AT_ROW + (((lane >> 2) ^ (k & 7)) << 4) + ((lane & 3) << 2));
        o0 = fmaf(pk, bf16_lo(vv), o0);
        o1 = fmaf(pk, bf16_hi(vv), o1);
      }
      sTr[tw * 64 + 2 * lane] = o0;
      sTr[tw * 64 + 2 * lane + 1] = o1;
      bar_sync_named(2, AT_TAIL_THREADS);
      if (tw == 0) {
        const float inv = 1.f / sum;
        o0 = (sTr[2 * lane] + sTr[64 + 2 * lane] + sTr[128 + 2 * lane]) * inv;
        o1 = (sTr[2 * lane + 1] + sTr[64 + 2 * lane + 1] + sTr[128 + 2 * lane + 1]) * inv;
        *reinterpret_cast<uint32_t*>(p.out + static_cast<long long>(row0 + qrow) * D + h * AT_DH + 2 * lane) = pack_bf16x2(o0, o1);
        if (lane == 0 && p.lse) p.lse[static_cast<long long>(row0 + qrow) * p.H + h] = (mx + log2f(sum)) * LN2;
      }
      bar_sync_named(2, AT_TAIL_THREADS);   // scratch is reused by the next tail row (and K / V / Q by the next item)
    }
    }  // items
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bar_arrive_named(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// 64-long dot product of a bf16 row in 128B-swizzled shared memory (row start address, row & 7) with a row held in
// registers as eight 16-byte chunks in logical order
__device__ __forceinline__ float dot64_swz(uint32_t row_addr, int r7, const uint4 (&w)[8]) {
  uint4 a[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c] = ld_shared_v4(row_addr + ((c ^ r7) << 4));
  float acc[4] = {0.f, 0.f, 0.f, 0.f};   // four chains: the dot product is latency- not throughput-bound
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    acc[0] = fmaf(bf16_lo(a[c].x), bf16_lo(w[c].x), fmaf(bf16_hi(a[c].x), bf16_hi(w[c].x), acc[0]));
    acc[1] = fmaf(bf16_lo(a[c].y), bf16_lo(w[c].y), fmaf(bf16_hi(a[c].y), bf16_hi(w[c].y), acc[1]));
    acc[2] = fmaf(bf16_lo(a[c].z), bf16_lo(w[c].z), fmaf(bf16_hi(a[c].z), bf16_hi(w[c].z), acc[2]));
    acc[3] = fmaf(bf16_lo(a[c].w), bf16_lo(w[c].w), fmaf(bf16_hi(a[c].w), bf16_hi(w[c].w), acc[3]));
  }
  return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}
__device__ __forceinline__ void load_row_swz(uint32_t base, int row, uint4 (&w)[8]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) w[c] = ld_shared_v4(base + row * AT_ROW + ((c ^ (row & 7)) << 4));
}
// 16 consecutive bf16 of a swizzled row (columns [16 part, 16 part + 16)) as floats
__device__ __forceinline__ void load_cols16_swz(uint32_t base, int row, int part, float (&f)[16]) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint4 a = ld_shared_v4(base + row * AT_ROW + (((2 * part + k) ^ (row & 7)) << 4));
    f[8 * k + 0] = bf16_lo(a.x); f[8 * k + 1] = bf16_hi(a.x); f[8 * k + 2] = bf16_lo(a.y); f[8 * k + 3] = bf16_hi(a.y);
    f[8 * k + 4] = bf16_lo(a.z); f[8 * k + 5] = bf16_hi(a.z); f[8 * k + 6] = bf16_lo(a.w); f[8 * k + 7] = bf16_hi(a.w);
  }
}

struct BwdParams {
  const __nv_bfloat16* out;
  const __nv_bfloat16* dout;
  const float* lse;
  const float* delta;   // [rows, H] rowsum(dO o O) from the out-projection dgrad epilogue; null = form it here from O
  __nv_bfloat16* dqkv;
  int row_base;
  int S, SP, H;
  int nt;     // 128-row tiles (queries and keys) that run on the tensor cores
  int ntail;  // trailing rows S - 128 nt (1..AT_TAIL) handled on the idle control warps, else 0
  int nbuf;  // TMEM score buffers (2 when nt <= 2)
  int prefetch;  // issue the next block's scores ahead of this block's accumulation (needs nbuf == 2)
  int lse_bulk;  // the sample's [S][H] lse block is 16-byte aligned: fetch it with one bulk copy
  int ahead;     // L2 prefetch of this CTA's next item on / off
  int nitems;    // (sample, head) pairs of the segment
  int pingpong;  // two softmax warp groups, one per 64-query half of a block (see the softmax section)
  int ntile;     // P^T / dS^T tile pairs in shared memory (2 when they fit: the softmax of block n + 1 then overlaps the
                 // accumulation MMAs of block n instead of waiting for them)
  int tl_cta;
  long long* tl;  // optional timeline buffer (tools/attn_timeline.py): CTA 0 records clock64() at its sync points
  float scale, scale_log2;
};

__global__ void __launch_bounds__(AT_BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tq128, const __grid_constant__ CUtensorMap tq16,
                   const __grid_constant__ CUtensorMap td128, const __grid_constant__ CUtensorMap td16,
                   const __grid_constant__ CUtensorMap tmdq, const __grid_constant__ CUtensorMap to128,
                   const __grid_constant__ CUtensorMap to16, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // Persistent like the forward: the CTA walks over (sample, head) items blockIdx.x, blockIdx.x + gridDim.x, ...;
  // step / block / key-tile counters run on across items so that every barrier keeps its phase sequence.
  const int S = p.S, SP = p.SP, nt = p.nt, nbuf = p.nbuf;
  const int D = p.H * AT_DH;
  // K and V are also read as 128-row A operands: the reads past row SP of K land in V, those of V in the
  // P^T tile — allocated memory whose content only reaches TMEM lanes that are masked below.
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + SP * AT_ROW;
  uint8_t* sK = sdO + SP * AT_ROW;
  uint8_t* sV = sK + SP * AT_ROW;
  uint8_t* sPt = sV + SP * AT_ROW;   // [128 kv][128 q] as two slabs of 64 q
  uint8_t* sdSt = sPt + 2 * AT_SLAB;
  // per-query statistics live in their own (static) arrays so that the compiler knows the score-tile stores
  // never alias them and can hoist their loads
  __shared__ __align__(16) float sLse[AT_STAT_N];     // lse * log2e
  __shared__ __align__(16) float sDelta[AT_STAT_N];   // rowsum(dO * O)
  // (a second P^T / dS^T pair follows when p.ntile == 2; staging for the epilogues always uses the first)
  constexpr uint32_t TILE_BYTES = 4 * AT_SLAB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdSt + 2 * AT_SLAB + (p.ntile - 1) * TILE_BYTES);
  uint64_t* bar_ld = bars + 0;      // Q, K, V have landed (the last of the two load groups)
  uint64_t* bar_ld0 = bars + 11;    // dO, O and the lse block have landed: delta can be formed while Q, K, V stream in
  uint64_t* bar_done = bars + 12;   // the softmax warps have finished the item (its operands, tiles and accumulators are free)
  uint64_t* bar_s = bars + 1;       // [2]
  uint64_t* bar_sfree = bars + 3;   // [2]
  uint64_t* bar_p = bars + 5;       // [2] one per 64-query half
  uint64_t* bar_tfree0 = bars + 7;   // tile pair t is free again (its accumulation MMAs have completed): t = 0 / 1
  uint64_t* bar_tfree1 = bars + 13;
  const int ntile = p.ntile;
  uint64_t* bar_acc = bars + 8;
  uint64_t* bar_accfree = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  // tail scratch (only with ntail > 0), each [AT_TAIL][SP] fp32: for tail key t and query x  P and dS (sPA, sDSA);
  // for tail query t and key x  P and dS (sPB, sDSB); then [9][64] partial sums of the tail rows' own outputs
  float* sPA = reinterpret_cast<float*>(bars + 16);
  float* sDSA = sPA + AT_TAIL * SP;
  float* sPB = sDSA + AT_TAIL * SP;
  float* sDSB = sPB + AT_TAIL * SP;
  float* sRed = sDSB + AT_TAIL * SP;
  const int M0 = 128 * nt;                       // rows [0, M0) go through the MMA path
  const int SPm = p.ntail > 0 ? M0 : SP;         // ... padded extent of that part

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto issue_loads = [&](int item) {
    const int h = item % p.H, sample = item / p.H;
    const int row0 = p.row_base + sample * S;
    // O (for delta) and this sample's [S][H] block of log-sum-exps are parked in the P^T / dS^T tile region,
    // which has no other use until the first score tile has been processed
    if (p.delta) {   // delta and lse come straight from global memory (one float each per query row): only dO to fetch
      mbar_expect_tx(bar_ld0, 1u * SP * AT_ROW);
      load_rows(sdO, &td128, &td16, bar_ld0, h * AT_DH, sample, SP);
    } else {
      const uint32_t lse_bytes = p.lse_bulk ? static_cast<uint32_t>(S) * p.H * 4u : 0u;
      mbar_expect_tx(bar_ld0, 2u * SP * AT_ROW + lse_bytes);
      load_rows(sdO, &td128, &td16, bar_ld0, h * AT_DH, sample, SP);
      load_rows(sPt, &to128, &to16, bar_ld0, h * AT_DH, sample, SP);
      if (p.lse_bulk) bulk_load_1d(sPt + SP * AT_ROW, p.lse + static_cast<long long>(row0) * p.H, lse_bytes, bar_ld0);
    }
    mbar_expect_tx(bar_ld, 3u * SP * AT_ROW);
    load_rows(sK, &tq128, &tq16, bar_ld, D + h * AT_DH, sample, SP);
    load_rows(sQ, &tq128, &tq16, bar_ld, h * AT_DH, sample, SP);
    load_rows(sV, &tq128, &tq16, bar_ld, 2 * D + h * AT_DH, sample, SP);
    const int nxt = item + static_cast<int>(gridDim.x);
    if (p.ahead > 0 && nxt < p.nitems) {
      const int h2 = nxt % p.H, s2 = nxt / p.H;
      prefetch_rows(&tq128, &tq16, D + h2 * AT_DH, s2, SP);
      prefetch_rows(&tq128, &tq16, h2 * AT_DH, s2, SP);
      prefetch_rows(&tq128, &tq16, 2 * D + h2 * AT_DH, s2, SP);
      prefetch_rows(&td128, &td16, h2 * AT_DH, s2, SP);
      if (!p.delta) prefetch_rows(&to128, &to16, h2 * AT_DH, s2, SP);
    }
  };
  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer: the first item's loads go out before
    // the set-up barrier (this thread owns the two load barriers), in flight while TMEM is allocated
    tma_prefetch_desc(&tq128);
    tma_prefetch_desc(&tq16);
    tma_prefetch_desc(&td128);
    tma_prefetch_desc(&td16);
    tma_prefetch_desc(&tmdq);
    tma_prefetch_desc(&to128);
    tma_prefetch_desc(&to16);
    mbar_init(bar_ld, 1);
    mbar_init(bar_ld0, 1);
    fence_barrier_init();
    issue_loads(blockIdx.x);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_sfree[i], p.pingpong ? AT_BWD_SM / 2 : AT_BWD_SM);   // one warp group per half-step with pingpong
      mbar_init(&bar_p[i], p.pingpong ? AT_BWD_SM / 2 : AT_BWD_SM);
    }
    mbar_init(bar_tfree0, 1);
    mbar_init(bar_tfree1, 1);
    mbar_init(bar_acc, 1);
    mbar_init(bar_accfree, AT_BWD_SM);
    mbar_init(bar_done, AT_BWD_SM);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t col_dv = nbuf * 128, col_dk = col_dv + 64, col_dq = col_dv + 128;

  if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-converged, see forward)
    constexpr uint32_t idesc_kt = make_idesc_bf16(128, 64, false, true);   // A K-major (P^T / dS^T), B MN-major (dO / Q)
    constexpr uint32_t idesc_mn = make_idesc_bf16(128, 64, true, true);    // A MN-major (dS), B MN-major (K)
    constexpr uint32_t ROW16 = AT_ROW / 16;                                // descriptor units per 128-byte row
    // descriptors built once; per MMA only a constant is added to the start-address field
    const uint64_t q_k = make_smem_desc_sw128(smem_u32(sQ), 0, 1024);      // Q  as K-major B   (scores)
    const uint64_t o_k = make_smem_desc_sw128(smem_u32(sdO), 0, 1024);     // dO as K-major B   (scores)
    const uint64_t k_k = make_smem_desc_sw128(smem_u32(sK), 0, 1024);      // K  as K-major A
    const uint64_t v_k = make_smem_desc_sw128(smem_u32(sV), 0, 1024);      // V  as K-major A
    const uint64_t q_mn = make_smem_desc_sw128(smem_u32(sQ), 8192, 1024);  // Q  as MN-major B  (dK)
    const uint64_t o_mn = make_smem_desc_sw128(smem_u32(sdO), 8192, 1024); // dO as MN-major B  (dV)
    const uint64_t k_mn = make_smem_desc_sw128(smem_u32(sK), 8192, 1024);  // K  as MN-major B  (dQ)
    const uint64_t pt_k = make_smem_desc_sw128(smem_u32(sPt), 0, 1024);    // P^T  as K-major A
    const uint64_t st_k = make_smem_desc_sw128(smem_u32(sdSt), 0, 1024);   // dS^T as K-major A
    const uint64_t st_mn = make_smem_desc_sw128(smem_u32(sdSt), AT_SLAB, 1024);  // dS^T tile as MN-major A (= dS)
    bool tl_on = false;
    int step = 0;          // score steps issued so far, over all items
    int jj = 0;            // key tiles finished so far, over all items
    uint32_t cnt_p[2] = {0, 0};
    int kit = 0;
    int gblk = 0;          // blocks finished so far, over all items (selects the P^T / dS^T tile pair)
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++kit) {
    mbar_wait(bar_ld0, kit & 1);
    mbar_wait(bar_ld, kit & 1);
    // the previous item's dQ epilogue has read its TMEM accumulators (and every other consumer is done)
    if (kit > 0) mbar_wait(bar_done, (kit - 1) & 1);
    tc_fence_after();
    TL(0, 0);
    // S^T = K_j Q_i^T and dP^T = V_j dO_i^T for 64-query half hh of block n = (j, i).  Steps are issued in order.
    auto halves_of = [&](int n) { return (min(128, SPm - 128 * (n % nt)) + 63) >> 6; };
    auto issue_step = [&](int n, int hh) {
      const int j = n / nt, i = n - j * nt;
      const int nq = min(128, SPm - 128 * i);
      const int b = step % nbuf, u = step / nbuf;
      if (u > 0) {
        mbar_wait(&bar_sfree[b], (u - 1) & 1);
        tc_fence_after();
      }
      const int nqh = min(64, nq - 64 * hh);
      const uint32_t idesc = make_idesc_bf16(128, nqh, false, false);
      const uint64_t kd = k_k + static_cast<uint64_t>(j * 128 * ROW16), vd = v_k + static_cast<uint64_t>(j * 128 * ROW16);
      const uint64_t qd = q_k + static_cast<uint64_t>((i * 128 + hh * 64) * ROW16);
      const uint64_t od = o_k + static_cast<uint64_t>((i * 128 + hh * 64) * ROW16);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + b * 128, kd + 2 * k, qd + 2 * k, idesc, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + b * 128 + 64, vd + 2 * k, od + 2 * k, idesc, k > 0 ? 1u : 0u);
        umma_commit(&bar_s[b]);
      }
      __syncwarp();
      ++step;
      TL(1, step);
    };
    const int nblk = nt * nt;
    tl_on = p.tl != nullptr && item == p.tl_cta && lane == 0;
    for (int hh = 0; hh < halves_of(0); ++hh) issue_step(0, hh);
    for (int n = 0; n < nblk; ++n) {
      const int j = n / nt, i = n - j * nt;
      const int nq = min(128, SPm - 128 * i), nkv = min(128, SPm - 128 * j);
      const int nh = (nq + 63) >> 6;
      // Score steps of the next block that can run ahead of this block's accumulation: all of them with two
      // TMEM score buffers, the first half with one (its buffer is free as soon as the softmax warps have read
      // the current scores; the second half needs them to have read the first, which at a key-tile boundary
      // only happens after the dK/dV epilogue that waits for this block's accumulation).
      int ahead = 0;
      if (p.prefetch && n + 1 < nblk) {
        ahead = (nbuf == 2) ? halves_of(n + 1) : 1;
        for (int hh = 0; hh < ahead; ++hh) issue_step(n + 1, hh);
      }
      for (int hh = 0; hh < nh; ++hh) {
        mbar_wait(&bar_p[hh], cnt_p[hh] & 1);
        ++cnt_p[hh];
      }
      tc_fence_after();
      TL(2, n);
      if (i == 0 && jj > 0) {   // dV / dK accumulators: the previous key tile's epilogue (of this or the last item) has read them
        mbar_wait(bar_accfree, (jj - 1) & 1);
        tc_fence_after();
      }
      const int kq = nq >> 4, kkv = nkv >> 4;
      const int tsel = gblk % ntile;
      const uint64_t toff = static_cast<uint64_t>(tsel) * (TILE_BYTES / 16);
      ++gblk;
      if (elect_one()) {
      // (fully unrolled with constant descriptor increments: the rolled loops cost ~60 cycles of address arithmetic
      // per MMA, 1.5 k cycles per block on the critical path between P^T / dS^T and the next block's tiles)
      {  // dV_j += P^T dO_i
        const uint64_t bb = o_mn + static_cast<uint64_t>(i * 128 * ROW16);
        const uint32_t d0 = tmem_base + col_dv;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          if (kk < kq)
            umma_f16_ss(d0, pt_k + toff + static_cast<uint64_t>((kk >> 2) * (AT_SLAB / 16) + (kk & 3) * 2),
                        bb + static_cast<uint64_t>(kk * 16 * ROW16), idesc_kt, (i > 0 || kk > 0) ? 1u : 0u);
      }
      {  // dK_j += dS^T Q_i
        const uint64_t bb = q_mn + static_cast<uint64_t>(i * 128 * ROW16);
        const uint32_t d0 = tmem_base + col_dk;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          if (kk < kq)
            umma_f16_ss(d0, st_k + toff + static_cast<uint64_t>((kk >> 2) * (AT_SLAB / 16) + (kk & 3) * 2),
                        bb + static_cast<uint64_t>(kk * 16 * ROW16), idesc_kt, (i > 0 || kk > 0) ? 1u : 0u);
      }
      {  // dQ_i += dS K_j   (dS^T tile read as an MN-major A operand)
        const uint64_t bb = k_mn + static_cast<uint64_t>(j * 128 * ROW16);
        const uint32_t d0 = tmem_base + col_dq + 64 * i;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          if (kk < kkv)
            umma_f16_ss(d0, st_mn + toff + static_cast<uint64_t>(kk * 16 * ROW16), bb + static_cast<uint64_t>(kk * 16 * ROW16), idesc_mn,
                        (j > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(tsel ? bar_tfree1 : bar_tfree0);
      if (i == nt - 1) umma_commit(bar_acc);
      }
      if (i == nt - 1) ++jj;
      __syncwarp();
      TL(3, n);
      if (n + 1 < nblk)
        for (int hh = ahead; hh < halves_of(n + 1); ++hh) issue_step(n + 1, hh);
    }
    }  // items
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax backward + epilogues
    const int quad = warp & 3;
    const int hf = (warp - 4) >> 2;  // which 16 of the 64 columns of a score half / of an output tile (0..3)
    const int r = quad * 32 + lane;  // key row inside the key tile (TMEM lane); query row in the dQ epilogue
    const int tid = threadIdx.x - 128;
    bool tl_on = false;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float sl2 = p.scale_log2;
    const int r7 = r & 7;
    const uint32_t pt_row = smem_u32(sPt) + r * AT_ROW, st_row = smem_u32(sdSt) + r * AT_ROW;
    int step = 0, blk = 0, jj = 0, kit = 0;   // counters over all items of this CTA (barrier phases)
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++kit) {
    const int h = item % p.H, sample = item / p.H;
    const int row0 = p.row_base + sample * S;
    tl_on = p.tl != nullptr && item == p.tl_cta && tid == 0;
    TL(4, 0);
    // delta = rowsum(dO * O) and lse in log2 units, both from shared memory, one query row per thread (the
    // 128B swizzle makes the row-per-lane reads conflict-free); the region O and the lse block are read from
    // becomes the P^T / dS^T tiles after the barrier below
    if (p.delta) {
      // one query row per thread: lse (log2 units) and delta from global memory, no wait on the operand loads
      const int q = tid;
      if (q < SP) {
        float ls = 0.f, dl = 0.f;
        if (q < S) {
          const long long o = static_cast<long long>(row0 + q) * p.H + h;
          ls = __ldg(p.lse + o) * LOG2E;
          dl = __ldg(p.delta + o);
        }
        sDelta[q] = dl;
        sLse[q] = ls;
      }
    } else {
      const float* lse_blk = reinterpret_cast<const float*>(sPt + SP * AT_ROW);   // [S][H]
      float lse_direct = 0.f;
      if (!p.lse_bulk && tid < S) lse_direct = p.lse[static_cast<long long>(row0 + tid) * p.H + h];
      TL(4, 2);
      mbar_wait(bar_ld0, kit & 1);
      TL(4, 3);
      const int q = tid;   // AT_BWD_SM >= AT_MAX_S: one query row per thread
      if (q < SP) {
        float dl = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int off = q * AT_ROW + ((c ^ (q & 7)) << 4);
          const uint4 a = *reinterpret_cast<const uint4*>(sPt + off);
          const uint4 g = *reinterpret_cast<const uint4*>(sdO + off);
          dl += bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x) + bf16_lo(a.y) * bf16_lo(g.y) +
                bf16_hi(a.y) * bf16_hi(g.y) + bf16_lo(a.z) * bf16_lo(g.z) + bf16_hi(a.z) * bf16_hi(g.z) +
                bf16_lo(a.w) * bf16_lo(g.w) + bf16_hi(a.w) * bf16_hi(g.w);
        }
        const float ls = (q < S) ? (p.lse_bulk ? lse_blk[q * p.H + h] : lse_direct) * LOG2E : 0.f;
        sDelta[q] = (q < S) ? dl : 0.f;
        sLse[q] = ls;
      }
    }
    TL(4, 4);
    bar_softmax_n<AT_BWD_SM>();
    if (p.ntail > 0) bar_arrive_named(3, AT_BWD_SM + AT_TAIL_THREADS);   // sLse / sDelta are complete
    TL(4, 1);
    bool store_pending = false;
    // Ping-pong: the 16 softmax warps form two groups of 8 (two per TMEM lane quadrant); group g owns the 64-query half
    // hh = g of every block and walks over its 32 columns per row in two chunks of 16.  In lockstep (all 16 warps on
    // every half) each thread paid the barrier round trip score-ready -> buffer-free -> tile-free -> fence -> P-ready
    // twice per block and the MUFU, FP32 and shared-memory-store phases of all warps coincided; with two groups a thread
    // pays it once per block and one group's exponentials overlap the other group's stores.
    const bool pp = p.pingpong != 0;
    const int grp = hf >> 1, sub = hf & 1;
    for (int j = 0; j < nt; ++j, ++jj) {
      if (pp && store_pending) {   // dK_j / dV_j of the previous key tile were staged in the P^T tile
        if (tid == 0) bulk_wait_read<0>();
        bar_softmax_n<AT_BWD_SM>();
        store_pending = false;
      }
      for (int i = 0; i < nt; ++i) {
        const int nq = min(128, SPm - 128 * i);
        const int nh = (nq + 63) >> 6;
        for (int hh = 0; hh < nh; ++hh, ++step) {
          if (pp && hh != grp) continue;
          const int b = step % nbuf, u = step / nbuf;
          mbar_wait(&bar_s[b], u & 1);
          tc_fence_after();
          TL(5, step);
          const int tsel = blk % ntile, tuse = blk / ntile;   // tile pair of this block and how often it has been used
          const uint32_t rowP = pt_row + tsel * TILE_BYTES + hh * AT_SLAB, rowS = st_row + tsel * TILE_BYTES + hh * AT_SLAB;
          const int nchunk = pp ? 2 : 1;
#pragma unroll 1
          for (int cc = 0; cc < nchunk; ++cc) {
            const int hfc = pp ? 2 * sub + cc : hf;   // which 16 of the 64 query columns of this half
            uint32_t sv[16], dv[16];
            tmem_ld_32x16(t_lane + b * 128 + 16 * hfc, sv);
            tmem_ld_32x16(t_lane + b * 128 + 64 + 16 * hfc, dv);
            tmem_ld_wait_dep16(sv);
            tmem_ld_wait_dep16(dv);
            if (cc == nchunk - 1) {
              tc_fence_before();
              mbar_arrive(&bar_sfree[b]);
            }
            if (cc == 0) {
              TL(6, step);
              if ((pp || hh == 0) && tuse > 0) mbar_wait(tsel ? bar_tfree1 : bar_tfree0, (tuse - 1) & 1);
              if (!pp && store_pending) {   // dK_j / dV_j of the previous key tile were staged in the P^T tile
                if (tid == 0) bulk_wait_read<0>();
                bar_softmax_n<AT_BWD_SM>();
                store_pending = false;
              }
              TL(7, step);
            }
            // No masking: query columns >= S have Q = dO = 0 (TMA zero fill), lse = delta = 0, hence P = 1 against a zero
            // dO row and dS = 0; key rows >= S only reach dK / dV rows that the TMA stores clip, and enter dQ against
            // zero-filled K rows (rows >= SP are never read by the dQ MMA).
            const int q0 = i * 128 + hh * 64 + 16 * hfc;
#pragma unroll
            for (int c = 0; c < 16; c += 8) {
              float pe[8], de[8];
              const float4 la = *reinterpret_cast<const float4*>(&sLse[q0 + c]);
              const float4 lb = *reinterpret_cast<const float4*>(&sLse[q0 + c + 4]);
              const float4 da = *reinterpret_cast<const float4*>(&sDelta[q0 + c]);
              const float4 db = *reinterpret_cast<const float4*>(&sDelta[q0 + c + 4]);
              const float l2[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
              const float dl[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                pe[e] = ex2(fmaf(__uint_as_float(sv[c + e]), sl2, -l2[e]));
                de[e] = pe[e] * (__uint_as_float(dv[c + e]) - dl[e]);
              }
              const uint32_t off = ((2 * hfc + (c >> 3)) ^ r7) << 4;
              st_shared_v4(rowP + off, pack_bf16x2(pe[0], pe[1]), pack_bf16x2(pe[2], pe[3]), pack_bf16x2(pe[4], pe[5]), pack_bf16x2(pe[6], pe[7]));
              st_shared_v4(rowS + off, pack_bf16x2(de[0], de[1]), pack_bf16x2(de[2], de[3]), pack_bf16x2(de[4], de[5]), pack_bf16x2(de[6], de[7]));
            }
          }
          TL(10, 2 + step);
          fence_proxy_async();
          mbar_arrive(&bar_p[hh]);
          TL(8, step);
        }
        ++blk;
      }
      // dK_j, dV_j: each half owns 32 of the 64 head-dim columns.  The two [128 x 64] bf16 tiles are staged in the
      // P^T tile (dead: bar_acc covers every MMA issued so far) and leave by TMA stores that clip rows >= S.
      mbar_wait(bar_acc, jj & 1);
      tc_fence_after();
      TL(9, j);
      {
        uint32_t a0[16], a1[16];
        tmem_ld_32x16(t_lane + col_dv + 16 * hf, a0);
        tmem_ld_32x16(t_lane + col_dk + 16 * hf, a1);
        tmem_ld_wait_dep16(a0);
        tmem_ld_wait_dep16(a1);
        tc_fence_before();
        mbar_arrive(bar_accfree);
        if (p.ntail > 0) {
          // tail queries t against this thread's key x:  dV[x] += P[t][x] dO[t],  dK[x] += dS[t][x] Q[t]
          if (j == 0) bar_sync_named(4, AT_BWD_SM + AT_TAIL_THREADS);   // the tail warps have filled the scratch
          const int x = j * 128 + r;
          for (int t = 0; t < p.ntail; ++t) {
            const float pb = sPB[t * SP + x], dsb = sDSB[t * SP + x];
            float f[16];
            load_cols16_swz(smem_u32(sdO), M0 + t, hf, f);
#pragma unroll
            for (int c = 0; c < 16; ++c) a0[c] = __float_as_uint(fmaf(pb, f[c], __uint_as_float(a0[c])));
            load_cols16_swz(smem_u32(sQ), M0 + t, hf, f);
#pragma unroll
            for (int c = 0; c < 16; ++c) a1[c] = __float_as_uint(fmaf(dsb, f[c], __uint_as_float(a1[c])));
          }
        }
        const float sc = p.scale;
#pragma unroll
        for (int c = 0; c < 16; c += 8) {
          const uint32_t off = ((2 * hf + (c >> 3)) ^ r7) << 4;
          st_shared_v4(pt_row + off,
              pack_bf16x2(__uint_as_float(a0[c]), __uint_as_float(a0[c + 1])), pack_bf16x2(__uint_as_float(a0[c + 2]), __uint_as_float(a0[c + 3])),
              pack_bf16x2(__uint_as_float(a0[c + 4]), __uint_as_float(a0[c + 5])), pack_bf16x2(__uint_as_float(a0[c + 6]), __uint_as_float(a0[c + 7])));
          st_shared_v4(pt_row + AT_SLAB + off,
              pack_bf16x2(__uint_as_float(a1[c]) * sc, __uint_as_float(a1[c + 1]) * sc), pack_bf16x2(__uint_as_float(a1[c + 2]) * sc, __uint_as_float(a1[c + 3]) * sc),
              pack_bf16x2(__uint_as_float(a1[c + 4]) * sc, __uint_as_float(a1[c + 5]) * sc), pack_bf16x2(__uint_as_float(a1[c + 6]) * sc, __uint_as_float(a1[c + 7]) * sc));
        }
        fence_proxy_async();
        bar_softmax_n<AT_BWD_SM>();
        if (tid == 0) {
          tma_store_3d(&tmdq, sPt, 2 * D + h * AT_DH, j * 128, sample);
          tma_store_3d(&tmdq, sPt + AT_SLAB, D + h * AT_DH, j * 128, sample);
          bulk_commit();
        }
        store_pending = true;
      }
    }
    TL(10, 0);
    // dQ_i (the last bar_acc phase covers every MMA issued): staged in the P^T / dS^T tiles once the dK / dV store
    // has drained, then one TMA store per query tile
    if (tid == 0) bulk_wait_read<0>();
    bar_softmax_n<AT_BWD_SM>();
    for (int i = 0; i < nt; ++i) {
      uint32_t a0[16];
      tmem_ld_32x16(t_lane + col_dq + 64 * i + 16 * hf, a0);
      tmem_ld_wait_dep16(a0);
      if (p.ntail > 0) {   // tail keys t against this thread's query x:  dQ[x] += dS[t][x] K[t]
        if (i == 0) bar_sync_named(5, AT_BWD_SM + AT_TAIL_THREADS);
        const int x = i * 128 + r;
        for (int t = 0; t < p.ntail; ++t) {
          const float dsa = sDSA[t * SP + x];
          float f[16];
          load_cols16_swz(smem_u32(sK), M0 + t, hf, f);
#pragma unroll
          for (int c = 0; c < 16; ++c) a0[c] = __float_as_uint(fmaf(dsa, f[c], __uint_as_float(a0[c])));
        }
      }
      const float sc = p.scale;
      const uint32_t stage = pt_row + i * AT_SLAB;   // slabs 0,1 of P^T, then slab 0 of dS^T (contiguous)
#pragma unroll
      for (int c = 0; c < 16; c += 8) {
        st_shared_v4(stage + (((2 * hf + (c >> 3)) ^ r7) << 4),
            pack_bf16x2(__uint_as_float(a0[c]) * sc, __uint_as_float(a0[c + 1]) * sc), pack_bf16x2(__uint_as_float(a0[c + 2]) * sc, __uint_as_float(a0[c + 3]) * sc),
            pack_bf16x2(__uint_as_float(a0[c + 4]) * sc, __uint_as_float(a0[c + 5]) * sc), pack_bf16x2(__uint_as_float(a0[c + 6]) * sc, __uint_as_float(a0[c + 7]) * sc));
      }
    }
    fence_proxy_async();
    bar_softmax_n<AT_BWD_SM>();
    if (tid == 0) {
      for (int i = 0; i < nt; ++i) tma_store_3d(&tmdq, sPt + i * AT_SLAB, h * AT_DH, i * 128, sample);
      bulk_commit();
      bulk_wait_read<0>();
    }
    // item finished: the staging tiles have been read by the stores, the accumulators by the epilogues, the operand
    // tiles by everybody -> the producer may load the next item, the MMA warp may overwrite TMEM
    tc_fence_before();
    bar_softmax_n<AT_BWD_SM>();
    mbar_arrive(bar_done);
    }  // items
  } else {
    // ------------------------------------------------------------------ warps 0, 2, 3: the producer thread loads the
    // following items; with a tail all 96 threads compute the tail rows on CUDA cores.
    // S = 128 nt + ntail (257-token decoders, the 260-token label-conditioned encoder): a third tile per dimension
    // would cost five more (j, i) blocks and the second TMEM score buffer.  For tail key t (row M0 + t) and every
    // query x:  P = exp2(q_x k_t sl2 - lse_x),  dS = P (dO_x v_t - delta_x);  for tail query t and every key x < M0:
    // P = exp2(q_t k_x sl2 - lse_t),  dS = P (dO_t v_x - delta_t).  The rows' own outputs are reduced here; what they
    // add to the rows of the main tiles is applied by the softmax warps in the dK / dV / dQ epilogues.
    const int tt = (warp == 0 ? 0 : warp - 1) * 32 + lane;   // 0..95
    const int tw = tt >> 5;
    const float sl2 = p.scale_log2;
    const uint32_t q_u = smem_u32(sQ), k_u = smem_u32(sK), v_u = smem_u32(sV), o_u = smem_u32(sdO);
    int kit = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++kit) {
    if (warp == 0 && lane == 0 && kit > 0) {
      mbar_wait(bar_done, (kit - 1) & 1);   // softmax warps (hence every MMA) and, in program order, the tail threads are done
      issue_loads(item);
    }
    __syncwarp();
    if (p.ntail == 0) continue;
    const int h = item % p.H, sample = item / p.H;
    const int row0 = p.row_base + sample * S;
    (void)sample;
    mbar_wait(bar_ld0, kit & 1);
    mbar_wait(bar_ld, kit & 1);
    bar_sync_named(3, AT_BWD_SM + AT_TAIL_THREADS);   // sLse / sDelta written by the softmax warps
    // tail queries first: the dK / dV epilogue of key tile 0 is the first consumer
    for (int t = 0; t < p.ntail; ++t) {
      const int rt = M0 + t;
      const float lse_t = sLse[rt], delta_t = sDelta[rt];
      uint4 w[8];
      load_row_swz(q_u, rt, w);
      for (int x = tt; x < M0; x += AT_TAIL_THREADS)
        sPB[t * SP + x] = ex2(fmaf(dot64_swz(k_u + x * AT_ROW, x & 7, w), sl2, -lse_t));
      load_row_swz(o_u, rt, w);
      for (int x = tt; x < M0; x += AT_TAIL_THREADS)
        sDSB[t * SP + x] = sPB[t * SP + x] * (dot64_swz(v_u + x * AT_ROW, x & 7, w) - delta_t);
    }
    bar_arrive_named(4, AT_BWD_SM + AT_TAIL_THREADS);   // sPB / sDSB complete (bar.arrive orders the writes before it)
    for (int t = 0; t < p.ntail; ++t) {
      const int rt = M0 + t;
      uint4 w[8];
      load_row_swz(k_u, rt, w);
      for (int x = tt; x < S; x += AT_TAIL_THREADS)
        sPA[t * SP + x] = ex2(fmaf(dot64_swz(q_u + x * AT_ROW, x & 7, w), sl2, -sLse[x]));
      load_row_swz(v_u, rt, w);
      for (int x = tt; x < S; x += AT_TAIL_THREADS)
        sDSA[t * SP + x] = sPA[t * SP + x] * (dot64_swz(o_u + x * AT_ROW, x & 7, w) - sDelta[x]);
    }
    bar_arrive_named(5, AT_BWD_SM + AT_TAIL_THREADS);   // sPA / sDSA complete: the dQ epilogue may read them
    bar_sync_named(2, AT_TAIL_THREADS);
    // the tail rows' own gradients: lane owns columns 2 lane, 2 lane + 1; warp tw takes rows tw, tw + 3, ...
    const uint32_t coff = ((lane & 3) << 2);
    for (int t = 0; t < p.ntail; ++t) {
      float v0 = 0.f, v1 = 0.f, k0 = 0.f, k1 = 0.f, g0 = 0.f, g1 = 0.f;
      for (int x = tw; x < S; x += 3) {
        const uint32_t off = x * AT_ROW + (((lane >> 2) ^ (x & 7)) << 4) + coff;
        const float pa = sPA[t * SP + x], dsa = sDSA[t * SP + x];
        uint32_t u = ld_shared_b32(o_u + off);
        v0 = fmaf(pa, bf16_lo(u), v0); v1 = fmaf(pa, bf16_hi(u), v1);          // dV[k_t] = sum_x P dO_x
        u = ld_shared_b32(q_u + off);
        k0 = fmaf(dsa, bf16_lo(u), k0); k1 = fmaf(dsa, bf16_hi(u), k1);        // dK[k_t] = sum_x dS Q_x
        if (x < M0) {
          const float dsb = sDSB[t * SP + x];
          u = ld_shared_b32(k_u + off);
          g0 = fmaf(dsb, bf16_lo(u), g0); g1 = fmaf(dsb, bf16_hi(u), g1);      // dQ[q_t] = sum_x dS K_x
        }
      }
      if (tw == 0) {   // ... plus the tail keys against this tail query
        for (int t2 = 0; t2 < p.ntail; ++t2) {
          const int x = M0 + t2;
          const float dsa = sDSA[t2 * SP + M0 + t];
          const uint32_t u = ld_shared_b32(k_u + x * AT_ROW + (((lane >> 2) ^ (x & 7)) << 4) + coff);
          g0 = fmaf(dsa, bf16_lo(u), g0); g1 = fmaf(dsa, bf16_hi(u), g1);
        }
      }
      float* red = sRed + tw * 192 + 2 * lane;
      red[0] = v0; red[1] = v1; red[64] = k0; red[65] = k1; red[128] = g0; red[129] = g1;
      bar_sync_named(2, AT_TAIL_THREADS);
      if (tw == 0) {
        const float* rr = sRed + 2 * lane;
        const float sc = p.scale;
        __nv_bfloat16* orow = p.dqkv + static_cast<long long>(row0 + M0 + t) * 3 * D + h * AT_DH + 2 * lane;
        *reinterpret_cast<uint32_t*>(orow + 2 * D) = pack_bf16x2(rr[0] + rr[192] + rr[384], rr[1] + rr[193] + rr[385]);
        *reinterpret_cast<uint32_t*>(orow + D) = pack_bf16x2((rr[64] + rr[256] + rr[448]) * sc, (rr[65] + rr[257] + rr[449]) * sc);
        *reinterpret_cast<uint32_t*>(orow) = pack_bf16x2((rr[128] + rr[320] + rr[512]) * sc, (rr[129] + rr[321] + rr[513]) * sc);
      }
      bar_sync_named(2, AT_TAIL_THREADS);
    }
    }  // items
  }

  if (p.tl != nullptr && static_cast<int>(blockIdx.x) == p.tl_cta % static_cast<int>(gridDim.x) && threadIdx.x == 128) p.tl[10 * 64 + 1] = clock64();
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct Segment {
  int n, S, row_base;
};
// rows of the last, partial 128-row box of a sample padded to 16 (16 when there is none: the map is then unused)
int tail_rows(int S) {
  const int t = ((S + 15) & ~15) % 128;
  return t ? t : 16;
}
// S = 128 * tiles + tail: a tail of 1..max_tail rows behind at least one full tile is computed on CUDA cores
void split_tail(int S, int max_tail, int* tiles, int* tail) {
  const int t = S % 128;
  if (S > 128 && t >= 1 && t <= max_tail) {
    *tiles = S / 128;
    *tail = t;
  } else {
    *tiles = (S + 127) / 128;
    *tail = 0;
  }
}
// Largest tail that goes to the control warps.  The three tail warps share their sub-partitions with the softmax
// warps, and in the persistent kernels the next item's loads wait for them (they read the operand tiles), so the
// tail work must stay well below the main loop.  Measured on B200 per layer (512 samples, 12 heads, persistent kernels):
//   257 tokens (every adaLN decoder):       forward 0.53 -> 0.35 ms, backward 1.05 -> 0.73 ms
//   258 tokens (decoder with a cond token): forward two rows pay (0.45 -> 0.41 ms); backward 1.05 without, 1.13 with
//   260 tokens (label-conditioned encoder): four rows are slower than the third tile in both directions
// hence the default limits: forward 2 rows, backward 1 row; the kernels handle up to AT_TAIL (tests raise the limits
// through umd_debug_attn_tail_limits).
int g_tail_limit[2] = {-1, -1};
int tail_limit(bool backward) {
  if (g_tail_limit[0] < 0) {
    const char* f = getenv("UMD_ATTN_TAIL_FWD_MAX");
    const char* b = getenv("UMD_ATTN_TAIL_BWD_MAX");
    g_tail_limit[0] = f ? atoi(f) : 2;
    g_tail_limit[1] = b ? atoi(b) : 1;
    for (int k = 0; k < 2; ++k) g_tail_limit[k] = g_tail_limit[k] < 0 ? 0 : (g_tail_limit[k] > AT_TAIL ? AT_TAIL : g_tail_limit[k]);
  }
  return g_tail_limit[backward ? 1 : 0];
}
int segments(const RowMap& rm, int nsamples, Segment (&seg)[2]) {
  int k = 0;
  const int n0 = rm.n0 < nsamples ? rm.n0 : nsamples;
  if (n0 > 0) seg[k++] = Segment{n0, rm.s0, 0};
  if (nsamples > n0) seg[k++] = Segment{nsamples - n0, rm.s1, rm.split_row};
  return k;
}

}  // namespace

bool attention_tc_supported(const RowMap& rm, int nsamples, int H, int Dh) {
  if (Dh != AT_DH || nsamples <= 0 || H <= 0) return false;
  Segment seg[2];
  const int ns = segments(rm, nsamples, seg);
  for (int k = 0; k < ns; ++k)
    if (seg[k].S < 1 || seg[k].S > AT_MAX_S) return false;
  return true;
}

int attention_fwd_tc(const AttnArgs& a, cudaStream_t st) {
  Segment seg[2];
  const int ns = segments(a.rm, a.nsamples, seg);
  const int D = a.H * AT_DH;
  static bool cfg = false;
  if (!cfg) {
    UMD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    UMD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    cfg = true;
  }
  for (int k = 0; k < ns; ++k) {
    const int n = seg[k].n, S = seg[k].S;
    // per-segment views [sample][row in sample][column]: loads zero-fill and stores clip rows >= S
    const __nv_bfloat16* qkv = a.qkv + static_cast<long long>(seg[k].row_base) * 3 * D;
    __nv_bfloat16* out = a.out + static_cast<long long>(seg[k].row_base) * D;
    CUtensorMap tm128, tm16, tmo;
    UMD_TRY(make_tmap_bf16(&tm128, qkv, 3 * D, S, n, 3 * D, static_cast<uint64_t>(S) * 3 * D, 128));
    UMD_TRY(make_tmap_bf16(&tm16, qkv, 3 * D, S, n, 3 * D, static_cast<uint64_t>(S) * 3 * D, tail_rows(S)));
    UMD_TRY(make_tmap_bf16(&tmo, out, D, S, n, D, static_cast<uint64_t>(S) * D, 128));
    FwdParams p;
    p.out = a.out; p.lse = a.lse; p.row_base = seg[k].row_base;
    p.S = S; p.SP = (p.S + 15) & ~15; p.H = a.H;
    split_tail(S, tail_limit(false), &p.nqt, &p.ntail);
    p.o_col = (p.SP + 31) & ~31;
    p.tmem_cols = (p.o_col + 128 <= 256) ? 256 : 512;   // scores + two output accumulators
    p.scale_log2 = a.scale * LOG2E;
    p.tl = g_attn_timeline;
    { const char* e = getenv("UMD_TL_CTA"); p.tl_cta = e ? atoi(e) : 0; }
    static int pf = -1;
    if (pf < 0) { const char* e = getenv("UMD_ATTN_L2PF"); pf = e ? atoi(e) : 1; }
    const int nslab = (p.SP + 63) / 64;
    const int qrows = p.nqt * 128 > p.SP ? p.nqt * 128 : p.SP;
    int smem = (qrows + 2 * p.SP) * AT_ROW + (nslab + 1) * AT_SLAB /*P + output staging*/ + 6144 /*row stats*/ + (AT_MAX_S + 208) * 4 /*tail scratch*/ + 256 + 1024;
    // a CTA that allocates 256 TMEM columns may share its SM with exactly one other
    if (p.tmem_cols == 256 && smem < 80 * 1024) smem = 80 * 1024;
    if (p.tmem_cols == 512 && smem < 120 * 1024) smem = 120 * 1024;
    const bool two = p.tmem_cols == 256 && smem <= 113 * 1024 && p.ntail == 0;
    p.ahead = pf;
    p.nitems = n * a.H;
    const int resident = sm_count() * (two ? 2 : 1);
    const int grid = p.nitems < resident ? p.nitems : resident;
    if (two)
      attn_fwd_tc_kernel<true><<<grid, 384, smem, st>>>(tm128, tm16, tmo, p);
    else
      attn_fwd_tc_kernel<false><<<grid, 640, smem, st>>>(tm128, tm16, tmo, p);
    ++g_launch_count;
    UMD_CHECK_CUDA(cudaGetLastError());
  }
  return UMD_OK;
}

int attention_bwd_tc(const AttnBwdArgs& a, cudaStream_t st) {
  Segment seg[2];
  const int ns = segments(a.rm, a.nsamples, seg);
  const int D = a.H * AT_DH;
  static bool cfg = false;
  if (!cfg) {
    UMD_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 223 * 1024));
    cfg = true;
  }
  for (int k = 0; k < ns; ++k) {
    const int n = seg[k].n, S = seg[k].S;
    const __nv_bfloat16* qkv = a.qkv + static_cast<long long>(seg[k].row_base) * 3 * D;
    const __nv_bfloat16* dout = a.dout + static_cast<long long>(seg[k].row_base) * D;
    __nv_bfloat16* dqkv = a.dqkv + static_cast<long long>(seg[k].row_base) * 3 * D;
    CUtensorMap tq128, tq16, td128, td16, tmdq;
    UMD_TRY(make_tmap_bf16(&tq128, qkv, 3 * D, S, n, 3 * D, static_cast<uint64_t>(S) * 3 * D, 128));
    UMD_TRY(make_tmap_bf16(&tq16, qkv, 3 * D, S, n, 3 * D, static_cast<uint64_t>(S) * 3 * D, tail_rows(S)));
    UMD_TRY(make_tmap_bf16(&td128, dout, D, S, n, D, static_cast<uint64_t>(S) * D, 128));
    UMD_TRY(make_tmap_bf16(&td16, dout, D, S, n, D, static_cast<uint64_t>(S) * D, tail_rows(S)));
    UMD_TRY(make_tmap_bf16(&tmdq, dqkv, 3 * D, S, n, 3 * D, static_cast<uint64_t>(S) * 3 * D, 128));
    UMD_REQUIRE(a.out || a.delta, "attention_bwd: either the forward output or delta = rowsum(dO o O) is required");
    // (with delta supplied the O maps are never used: point them at dO so that they stay valid descriptors)
    const __nv_bfloat16* outp = a.delta ? dout : a.out + static_cast<long long>(seg[k].row_base) * D;
    CUtensorMap to128, to16;
    UMD_TRY(make_tmap_bf16(&to128, outp, D, S, n, D, static_cast<uint64_t>(S) * D, 128));
    UMD_TRY(make_tmap_bf16(&to16, outp, D, S, n, D, static_cast<uint64_t>(S) * D, tail_rows(S)));
    BwdParams p;
    p.lse_bulk = ((S * a.H * 4) % 16 == 0) && ((static_cast<long long>(seg[k].row_base) * a.H * 4) % 16 == 0) &&
                 ((reinterpret_cast<uintptr_t>(a.lse) & 15) == 0) && (((S + 15) & ~15) * AT_ROW + S * a.H * 4 <= 4 * AT_SLAB);
    p.out = a.out; p.dout = a.dout; p.lse = a.lse; p.delta = a.delta; p.dqkv = a.dqkv; p.row_base = seg[k].row_base;
    p.S = S; p.SP = (p.S + 15) & ~15; p.H = a.H;
    split_tail(S, tail_limit(true), &p.nt, &p.ntail);
    p.nbuf = p.nt <= 2 ? 2 : 1;
    { const char* e = getenv("UMD_ATTN_PREFETCH"); p.prefetch = e ? atoi(e) != 0 : 1; }
    p.scale = a.scale; p.scale_log2 = a.scale * LOG2E;
    p.tl = g_attn_timeline;
    { const char* e = getenv("UMD_TL_CTA"); p.tl_cta = e ? atoi(e) : 0; }
    static int pf = -1;
    if (pf < 0) { const char* e = getenv("UMD_ATTN_L2PF"); pf = e ? atoi(e) : 1; }
    p.ahead = pf;
    p.nitems = n * a.H;
    static int pingpong = -1;
    if (pingpong < 0) { const char* e = getenv("UMD_ATTN_PINGPONG"); pingpong = e ? atoi(e) : 1; }
    // (with one TMEM score buffer the two groups would have to share one score-ready barrier and skip each other's
    // phases, which a parity wait cannot express: lockstep there)
    p.pingpong = (pingpong && p.nbuf == 2) ? 1 : 0;
    static int tiles2 = -1;
    if (tiles2 < 0) { const char* e = getenv("UMD_ATTN_BWD_TILES"); tiles2 = e ? atoi(e) : 2; }
    int smem = 4 * p.SP * AT_ROW + 4 * AT_SLAB + 256 + 1024;   // + 3 KB of static shared memory
    if (p.ntail > 0) smem += (4 * AT_TAIL * p.SP + 9 * 64) * 4;
    p.ntile = (tiles2 >= 2 && smem + 4 * AT_SLAB <= 220 * 1024) ? 2 : 1;
    smem += (p.ntile - 1) * 4 * AT_SLAB;
    if (smem < 120 * 1024) smem = 120 * 1024;  // the kernel owns all 512 TMEM columns: one CTA per SM
    const int grid = p.nitems < sm_count() ? p.nitems : sm_count();
    attn_bwd_tc_kernel<<<grid, AT_BWD_THREADS, smem, st>>>(tq128, tq16, td128, td16, tmdq, to128, to16, p);
    ++g_launch_count;
    UMD_CHECK_CUDA(cudaGetLastError());
  }
  return UMD_OK;
}

}  // namespace umd

// test hook: largest tail (0..4) routed to the control warps, forward / backward; negative = back to the defaults
extern "C" void umd_debug_attn_tail_limits(int fwd_max, int bwd_max) {
  if (fwd_max < 0 || bwd_max < 0) {
    umd::g_tail_limit[0] = umd::g_tail_limit[1] = -1;
    return;
  }
  umd::g_tail_limit[0] = fwd_max > umd::AT_TAIL ? umd::AT_TAIL : fwd_max;
  umd::g_tail_limit[1] = bwd_max > umd::AT_TAIL ? umd::AT_TAIL : bwd_max;
}

// debug aid (tools/attn_timeline.py): device buffer of 11 x 64 clock64() stamps written by CTA 0 of the backward kernel
extern "C" void umd_debug_attn_timeline(long long* device_buf) { umd::g_attn_timeline = device_buf; }
