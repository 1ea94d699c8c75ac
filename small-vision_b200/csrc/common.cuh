// Shared host/device helpers for the UMD hot-path library.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/umd_b200.h"

namespace umd {

// ---------------------------------------------------------------------------
// Error plumbing: every extern "C" entry returns 0 on success; the message of the
// last failure on the calling thread is available through umd_last_error().
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define UMD_CHECK_CUDA(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::umd::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return UMD_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define UMD_REQUIRE(cond, ...)         \
  do {                                 \
    if (!(cond)) {                     \
      ::umd::set_error(__VA_ARGS__);   \
      return UMD_ERR_INVALID;          \
    }                                  \
  } while (0)

#define UMD_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

// ---------------------------------------------------------------------------
// Two-segment ragged batch: rows [0, split_row) belong to samples 0..n0-1 with s0
// rows each (noise branch); the rest to samples n0.. with s1 rows each (clean / MAE
// branch).  A uniform batch sets split_row = total rows, s1 = 1.
// ---------------------------------------------------------------------------
struct RowMap {
  int split_row;
  int s0;
  int s1;
  int n0;
};

__host__ __device__ __forceinline__ int sample_of(const RowMap& m, int row) {
  return row < m.split_row ? row / m.s0 : m.n0 + (row - m.split_row) / m.s1;
}
// token index of `row` inside its sample
__host__ __device__ __forceinline__ int token_of(const RowMap& m, int row) {
  return row < m.split_row ? row % m.s0 : (row - m.split_row) % m.s1;
}
__host__ __device__ __forceinline__ int row_of(const RowMap& m, int sample, int tok) {
  return sample < m.n0 ? sample * m.s0 + tok : m.split_row + (sample - m.n0) * m.s1 + tok;
}
__host__ __device__ __forceinline__ int seq_of(const RowMap& m, int sample) { return sample < m.n0 ? m.s0 : m.s1; }

inline RowMap uniform_rowmap(int n, int s) {
  RowMap m;
  m.split_row = n * s;
  m.s0 = s;
  m.s1 = 1;
  m.n0 = n;
  return m;
}
inline RowMap ragged_rowmap(int n0, int s0, int n1, int s1) {
  RowMap m;
  m.split_row = n0 * s0;
  m.s0 = s0 > 0 ? s0 : 1;
  m.s1 = (n1 > 0 && s1 > 0) ? s1 : 1;
  m.n0 = n0;
  return m;
}

// ---------------------------------------------------------------------------
// TMA descriptor encode (driver entry point resolved at run time so that the
// library links against cudart only and still loads on a machine without a GPU).
// ---------------------------------------------------------------------------
// bf16 tensor viewed as [batch][outer][inner] with `inner` contiguous; box = {64, box_outer, 1},
// 128-byte swizzle, zero fill out of bounds.
int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t batch,
                   uint64_t ld_elems, uint64_t batch_stride_elems, uint32_t box_outer);

int sm_count();

// ---------------------------------------------------------------------------
// Optional per-category device timing (bench.py's live roofline): when enabled, the leaf launchers
// bracket their kernels with CUDA events on the launching stream and account the algorithmic work
// (FLOPs for tensor-bound categories, bytes for HBM-bound ones).  Off by default; costs nothing then.
// ---------------------------------------------------------------------------
enum ProfCat {
  PC_GEMM = 0,      // forward / dgrad GEMMs (activations as the A operand)
  PC_GEMM_WGRAD,    // weight-gradient GEMMs (both operands MN-major, split-K)
  PC_ATTN_FWD,
  PC_ATTN_BWD,
  PC_LN_FWD,
  PC_LN_BWD,
  PC_GATE_BWD,
  PC_COLSUM,
  PC_OPTIMIZER,
  PC_COUNT
};
extern bool g_prof_on;
void prof_open(int cat, double work, cudaStream_t st);
void prof_close(cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(int cat, double work, cudaStream_t s) : st(s), on(g_prof_on) {
    if (on) prof_open(cat, work, s);
  }
  ~ProfScope() {
    if (on) prof_close(st);
  }
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace umd
