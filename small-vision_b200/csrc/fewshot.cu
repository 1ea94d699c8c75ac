// Few-shot ridge probe on `pre_logits` (big_vision/evaluators/fewshot_lsr.py:43-116).
//
// The reference whitens the support features, appends a constant bias feature (100), regresses one-hot targets in
// {-1, +1} with an L2 penalty and scores the query set by argmax.  It solves the normal equations through an
// eigendecomposition so that several penalties share one factorisation; the ridge solution itself is
//   N >= D:  W = (X^T X + l2 I)^-1 X^T Y          N < D:  W = X^T (X X^T + l2 I)^-1 Y
// which is what is computed here: Gram matrices and X^T Y in fp32 on the CUDA cores (all kernels HBM/L2 or FMA bound,
// nothing GEMM-heavy enough for the tensor pipe at D = 769), the symmetric positive-definite solve by a Cholesky
// factorisation in fp64 (the bias feature puts 1e9 on the diagonal against l2 = 1024: condition ~1e6).
#include "common.cuh"

namespace umd {
extern long long g_launch_count;

#define FS_LAUNCH_CHECK()                \
  do {                                   \
    ++g_launch_count;                    \
    UMD_CHECK_CUDA(cudaGetLastError());  \
  } while (0)

// -----------------------------------------------------------------------------------------
// Column statistics: mean and population standard deviation (+1e-5) of x[n, d]   (fewshot_lsr.py:46-47)
// One CTA per 32 columns, 32 x 8 threads: coalesced 128-byte row segments, two passes (mean, then centred squares).
// -----------------------------------------------------------------------------------------
__global__ void fs_stats_kernel(const float* __restrict__ x, int n, int d, float* __restrict__ mean,
                                float* __restrict__ stdv) {
  __shared__ double red[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const bool live = col < d;
  double s = 0.0;
  for (int r = threadIdx.y; r < n; r += 8) s += live ? static_cast<double>(x[static_cast<long long>(r) * d + col]) : 0.0;
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  double m = 0.0;
  for (int k = 0; k < 8; ++k) m += red[k][threadIdx.x];
  m /= n;
  __syncthreads();
  double q = 0.0;
  for (int r = threadIdx.y; r < n; r += 8) {
    const double v = live ? static_cast<double>(x[static_cast<long long>(r) * d + col]) - m : 0.0;
    q += v * v;
  }
  red[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && live) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    mean[col] = static_cast<float>(m);
    stdv[col] = static_cast<float>(sqrt(t / n)) + 1e-5f;
  }
}

// out[n, d + 1] = [(x - mean) / std | 100]                                        (fewshot_lsr.py:48-51,99-100)
__global__ void fs_whiten_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ stdv,
                                 long long n, int d, float bias_constant, float* __restrict__ out) {
  const long long total = n * (d + 1);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / (d + 1);
    const int c = static_cast<int>(i - r * (d + 1));
    out[i] = c < d ? (x[r * d + c] - mean[c]) / stdv[c] : bias_constant;
  }
}

// rhs[dim, c] = X^T Y with Y = 2 onehot(y) - 1 (fewshot_lsr.py:54): 2 * (sum of the rows of class c) - (sum of all rows).
// Pass 1 scatters rows into class sums (fp32 atomics: every class receives `shots` rows, so contention is low), pass 2
// adds the class sums up to the total row (one thread per feature, coalesced), pass 3 finishes.  Rows whose label is
// outside [0, c) only count towards the total (their one-hot row is all zero), like jax.nn.one_hot.
__global__ void fs_class_sums_kernel(const float* __restrict__ xw, const int* __restrict__ y, int n, int dim, int c,
                                     float* __restrict__ sums /*[c + 2, dim], zeroed; row c = total, row c + 1 = unlabelled*/) {
  const int r = blockIdx.x;
  int cls = y[r];
  if (cls < 0 || cls >= c) cls = c + 1;
  const float* row = xw + static_cast<long long>(r) * dim;
  for (int k = threadIdx.x; k < dim; k += blockDim.x) atomicAdd(&sums[static_cast<long long>(cls) * dim + k], row[k]);
}
__global__ void fs_total_row_kernel(float* __restrict__ sums, int dim, int c) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= dim) return;
  float t = sums[static_cast<long long>(c + 1) * dim + k];
  for (int cls = 0; cls < c; ++cls) t += sums[static_cast<long long>(cls) * dim + k];
  sums[static_cast<long long>(c) * dim + k] = t;
}
__global__ void fs_rhs_from_sums_kernel(const float* __restrict__ sums, int dim, int c, float* __restrict__ rhs /*[dim, c]*/) {
  const long long total = static_cast<long long>(dim) * c;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i / c), cls = static_cast<int>(i - static_cast<long long>(k) * c);
    rhs[i] = 2.f * sums[static_cast<long long>(cls) * dim + k] - sums[static_cast<long long>(c) * dim + k];
  }
}
// Y[n, c] = 2 onehot(y) - 1 (the N < D branch needs the targets themselves, fewshot_lsr.py:88)
__global__ void fs_targets_kernel(const int* __restrict__ y, long long n, int c, float* __restrict__ out) {
  const long long total = n * c;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / c;
    out[i] = (y[r] == static_cast<int>(i - r * c)) ? 1.f : -1.f;
  }
}

// -----------------------------------------------------------------------------------------
// fp32 matrix product on the CUDA cores with arbitrary element strides (so that X^T X, X X^T, X^T Z and X W are the
// same kernel):  C[m, n] = sum_k A(m, k) B(k, n),  A(m, k) = A[m * a_rs + k * a_cs],  B(k, n) = B[k * b_rs + n * b_cs].
// 128 x 128 tile per CTA, 256 threads, 8 x 8 outputs per thread, K staged 16 at a time through shared memory
// (stored k-major so that the inner product reads are conflict-free broadcasts / consecutive words).  Optional
// split over K (gridDim.z) with atomic accumulation for the tall-skinny Gram products (K = support size).
// -----------------------------------------------------------------------------------------
constexpr int FS_TM = 128, FS_TN = 128, FS_TK = 16;
__global__ void __launch_bounds__(256)
fs_matmul_kernel(const float* __restrict__ A, long long a_rs, long long a_cs, const float* __restrict__ B, long long b_rs,
                 long long b_cs, float* __restrict__ C, long long ldc, int M, int N, int K, int k_per_split, int atomic) {
  __shared__ float As[FS_TK][FS_TM + 4];
  __shared__ float Bs[FS_TK][FS_TN + 4];
  const int m0 = blockIdx.y * FS_TM, n0 = blockIdx.x * FS_TN;
  const int kb = blockIdx.z * k_per_split, ke = min(K, kb + k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // 16 x 16 threads, each 8 x 8 (strided by 16)
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  // loader mapping: which of the two strides is unit decides which index runs fastest across threads
  const bool a_k_fast = (a_cs == 1), b_n_fast = (b_cs == 1);
  for (int k0 = kb; k0 < ke; k0 += FS_TK) {
    for (int e = threadIdx.x; e < FS_TM * FS_TK; e += 256) {
      int mm, kk;
      if (a_k_fast) { kk = e % FS_TK; mm = e / FS_TK; } else { mm = e % FS_TM; kk = e / FS_TM; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < ke) ? A[gm * a_rs + gk * a_cs] : 0.f;
    }
    for (int e = threadIdx.x; e < FS_TN * FS_TK; e += 256) {
      int nn, kk;
      if (b_n_fast) { nn = e % FS_TN; kk = e / FS_TN; } else { kk = e % FS_TK; nn = e / FS_TK; }
      const int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < ke) ? B[gk * b_rs + gn * b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < FS_TK; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty + 16 * i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + tx + 16 * j;
      if (gn >= N) continue;
      float* dst = C + gm * ldc + gn;
      if (atomic) atomicAdd(dst, acc[i][j]); else *dst = acc[i][j];
    }
  }
}

// -----------------------------------------------------------------------------------------
// Ridge solve: (G + l2 I) Z = R for symmetric positive semi-definite G [n, n] (fp32 in) and R [n, c] (fp32 in, Z out).
// fp64 work copy (L2-resident for n <= ~1k); blocked Cholesky factorisation, then forward / backward substitution.
// -----------------------------------------------------------------------------------------
__global__ void fs_load_system_kernel(const float* __restrict__ G, float l2, int n, double* __restrict__ L) {
  const long long total = static_cast<long long>(n) * n;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / n), c = static_cast<int>(i - static_cast<long long>(r) * n);
    // symmetrise (atomic split-K accumulation orders differ between the two triangles)
    const double v = 0.5 * (static_cast<double>(G[i]) + static_cast<double>(G[static_cast<long long>(c) * n + r]));
    L[i] = v + (r == c ? static_cast<double>(l2) : 0.0);
  }
}
// Blocked right-looking Cholesky, NB = 32, three kernels per block column (the launches are the grid-wide barriers):
//   diag  : one CTA factorises the 32 x 32 diagonal block in shared memory;
//   panel : rows below it, X = A L_kk^-T, one row per thread out of a padded shared tile;
//   update: every lower-triangular 32 x 32 tile pair of the trailing matrix, C -= P_i P_j^T (K = 32 from shared memory).
// A last pass mirrors L into the upper triangle so that the back substitution reads rows.
constexpr int FS_NB = 32;
__global__ void __launch_bounds__(1024) fs_chol_diag_kernel(double* __restrict__ L, int n, int kb, int* __restrict__ status) {
  __shared__ double a[FS_NB][FS_NB + 1];
  const int r = threadIdx.x >> 5, c = threadIdx.x & 31;
  const int nb = min(FS_NB, n - kb);
  a[r][c] = (r < nb && c < nb) ? L[static_cast<long long>(kb + r) * n + kb + c] : (r == c ? 1.0 : 0.0);
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    if (r == j && c == j) {
      const double d = a[j][j];
      if (!(d > 0.0)) { *status = kb + j + 1; a[j][j] = 1.0; } else a[j][j] = sqrt(d);
    }
    __syncthreads();
    if (c == j && r > j) a[r][j] /= a[j][j];
    __syncthreads();
    if (c > j && r >= c) a[r][c] -= a[r][j] * a[c][j];
    __syncthreads();
  }
  if (r < nb && c <= r) L[static_cast<long long>(kb + r) * n + kb + c] = a[r][c];
}
__global__ void __launch_bounds__(128) fs_chol_panel_kernel(double* __restrict__ L, int n, int kb) {
  __shared__ double lkk[FS_NB][FS_NB + 1];
  __shared__ double rows[128][FS_NB + 1];
  const int nb = min(FS_NB, n - kb);
  const int row0 = kb + nb + blockIdx.x * 128;
  for (int e = threadIdx.x; e < FS_NB * FS_NB; e += 128) {
    const int r = e >> 5, c = e & 31;
    lkk[r][c] = (r < nb && c <= r) ? L[static_cast<long long>(kb + r) * n + kb + c] : (r == c ? 1.0 : 0.0);
  }
  for (int e = threadIdx.x; e < 128 * FS_NB; e += 128) {
    const int r = e >> 5, c = e & 31;
    rows[r][c] = (row0 + r < n && c < nb) ? L[static_cast<long long>(row0 + r) * n + kb + c] : 0.0;
  }
  __syncthreads();
  double* x = rows[threadIdx.x];
  for (int c = 0; c < nb; ++c) {
    double s = x[c];
    for (int m = 0; m < c; ++m) s -= x[m] * lkk[c][m];
    x[c] = s / lkk[c][c];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 128 * FS_NB; e += 128) {
    const int r = e >> 5, c = e & 31;
    if (row0 + r < n && c < nb) L[static_cast<long long>(row0 + r) * n + kb + c] = rows[r][c];
  }
}
__global__ void __launch_bounds__(1024) fs_chol_update_kernel(double* __restrict__ L, int n, int kb) {
  if (blockIdx.x > blockIdx.y) return;             // lower-triangular tile pairs only (x = column block, y = row block)
  __shared__ double pa[FS_NB][FS_NB + 1], pb[FS_NB][FS_NB + 1];
  const int r = threadIdx.x >> 5, c = threadIdx.x & 31;
  const int base = kb + FS_NB;                      // only called while a full block column precedes the trailing matrix
  const int gi = base + blockIdx.y * FS_NB + r, gj = base + blockIdx.x * FS_NB + r;
  pa[r][c] = gi < n ? L[static_cast<long long>(gi) * n + kb + c] : 0.0;
  pb[r][c] = gj < n ? L[static_cast<long long>(gj) * n + kb + c] : 0.0;
  __syncthreads();
  const int oi = base + blockIdx.y * FS_NB + r, oj = base + blockIdx.x * FS_NB + c;
  if (oi >= n || oj > oi) return;
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < FS_NB; ++k) s += pa[r][k] * pb[c][k];
  L[static_cast<long long>(oi) * n + oj] -= s;
}
__global__ void fs_mirror_kernel(double* __restrict__ L, int n) {
  const long long total = static_cast<long long>(n) * n;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / n), c = static_cast<int>(i - static_cast<long long>(r) * n);
    if (c > r) L[i] = L[static_cast<long long>(c) * n + r];
  }
}
// Forward then backward substitution.  A CTA (4 warps) owns `cols` right-hand sides (32, 16 or 8: whatever keeps the
// fp64 solution block W[n][cols] in shared memory); the dot product of row i is split over 4 * 32 / cols thread groups
// (k strided), reduced by shuffles inside a warp and through shared memory across warps.  Rows of L (and of L^T, kept
// in the upper triangle by the factorisation) are staged four at a time into shared memory with coalesced loads, so the
// L2 latency is paid once per four rows instead of once per element.
constexpr int FS_RING = 4;
__global__ void __launch_bounds__(128)
fs_trisolve_kernel(const double* __restrict__ L, int n, int c, int cols, const float* __restrict__ R, float* __restrict__ Z) {
  extern __shared__ double fs_sm[];
  double* W = fs_sm;                                        // [n][cols]
  double* ring = W + static_cast<size_t>(n) * cols;         // [FS_RING][n]
  double* part = ring + static_cast<size_t>(FS_RING) * n;   // [4][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = lane % cols, sub = lane / cols, per_warp = 32 / cols;
  const int kg = warp * per_warp + sub, KG = 4 * per_warp;
  const int t = blockIdx.x * cols + col;
  const bool live = t < c;
  for (int pass = 0; pass < 2; ++pass) {
    for (int step0 = 0; step0 < n; step0 += FS_RING) {
      const int rows_here = min(FS_RING, n - step0);
      for (int rr = 0; rr < rows_here; ++rr) {              // stage rows (diagonal element included)
        const int i = pass == 0 ? step0 + rr : n - 1 - (step0 + rr);
        const double* Li = L + static_cast<long long>(i) * n;
        const int lo = pass == 0 ? 0 : i, hi = pass == 0 ? i + 1 : n;
        for (int k = lo + threadIdx.x; k < hi; k += 128) ring[rr * n + k] = Li[k];
      }
      __syncthreads();
      for (int rr = 0; rr < rows_here; ++rr) {
        const int i = pass == 0 ? step0 + rr : n - 1 - (step0 + rr);
        const double* Li = ring + rr * n;
        const int k_lo = pass == 0 ? 0 : i + 1, k_hi = pass == 0 ? i : n;
        double s0 = 0.0, s1 = 0.0;
        int k = k_lo + kg;
        for (; k + KG < k_hi; k += 2 * KG) {
          s0 += Li[k] * W[static_cast<size_t>(k) * cols + col];
          s1 += Li[k + KG] * W[static_cast<size_t>(k + KG) * cols + col];
        }
        if (k < k_hi) s0 += Li[k] * W[static_cast<size_t>(k) * cols + col];
        double s = s0 + s1;
        for (int o = cols; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (sub == 0) part[warp * 32 + col] = s;
        __syncthreads();
        if (warp == 0 && sub == 0) {
          const double r = pass == 0 ? (live ? static_cast<double>(R[static_cast<long long>(i) * c + t]) : 0.0)
                                     : W[static_cast<size_t>(i) * cols + col];
          const double w = (r - (part[col] + part[32 + col] + part[64 + col] + part[96 + col])) / Li[i];
          W[static_cast<size_t>(i) * cols + col] = w;
          if (pass == 1 && live) Z[static_cast<long long>(i) * c + t] = static_cast<float>(w);
        }
        __syncthreads();
      }
    }
  }
}

// preds = argmax(scores, axis=1) (first maximum, like jnp.argmax); counts preds == labels  (fewshot_lsr.py:111-112)
__global__ void fs_accuracy_kernel(const float* __restrict__ scores, const int* __restrict__ labels, int n, int c,
                                   int* __restrict__ preds_or_null, int* __restrict__ correct) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* row = scores + static_cast<long long>(warp) * c;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int k = lane; k < c; k += 32) {
    const float v = row[k];
    if (v > best || (v != v && best == best)) { best = v; arg = k; }   // NaN wins like in jnp.argmax
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    const bool take = (ob > best) || (ob == best && oa < arg) || (ob != ob && (best == best || oa < arg));
    if (take) { best = ob; arg = oa; }
  }
  if (lane == 0) {
    if (arg == 0x7fffffff) arg = 0;
    if (preds_or_null) preds_or_null[warp] = arg;
    if (arg == labels[warp]) atomicAdd(correct, 1);
  }
}

static int grid_for(long long total, int threads) {
  long long g = ceil_div_ll(total, threads);
  const long long cap = static_cast<long long>(sm_count()) * 8;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}
}  // namespace umd

using namespace umd;

extern "C" int umd_fewshot_stats(const float* x, int n, int d, float* mean, float* std_plus_eps, umd_stream_t stream) {
  UMD_REQUIRE(x && mean && std_plus_eps && n > 0 && d > 0, "umd_fewshot_stats: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  fs_stats_kernel<<<static_cast<int>(ceil_div_ll(d, 32)), dim3(32, 8), 0, st>>>(x, n, d, mean, std_plus_eps);
  FS_LAUNCH_CHECK();
  return UMD_OK;
}
extern "C" int umd_fewshot_whiten(const float* x, const float* mean, const float* std_plus_eps, int n, int d,
                                  float bias_constant, float* out, umd_stream_t stream) {
  UMD_REQUIRE(x && mean && std_plus_eps && out && n >= 0 && d > 0, "umd_fewshot_whiten: bad argument");
  if (n == 0) return UMD_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(n) * (d + 1);
  fs_whiten_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, mean, std_plus_eps, n, d, bias_constant, out);
  FS_LAUNCH_CHECK();
  return UMD_OK;
}
extern "C" int umd_fewshot_xty(const float* xw, const int* y, int n, int dim, int num_classes, float* sums_scratch,
                               float* rhs, umd_stream_t stream) {
  UMD_REQUIRE(xw && y && sums_scratch && rhs && n > 0 && dim > 0 && num_classes > 0, "umd_fewshot_xty: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UMD_CHECK_CUDA(cudaMemsetAsync(sums_scratch, 0, sizeof(float) * (num_classes + 2ll) * dim, st));
  fs_class_sums_kernel<<<n, 256, 0, st>>>(xw, y, n, dim, num_classes, sums_scratch);
  FS_LAUNCH_CHECK();
  fs_total_row_kernel<<<static_cast<int>(ceil_div_ll(dim, 128)), 128, 0, st>>>(sums_scratch, dim, num_classes);
  FS_LAUNCH_CHECK();
  fs_rhs_from_sums_kernel<<<grid_for(static_cast<long long>(dim) * num_classes, 256), 256, 0, st>>>(sums_scratch, dim,
                                                                                                    num_classes, rhs);
  FS_LAUNCH_CHECK();
  return UMD_OK;
}
extern "C" int umd_fewshot_targets(const int* y, int n, int num_classes, float* out, umd_stream_t stream) {
  UMD_REQUIRE(y && out && n > 0 && num_classes > 0, "umd_fewshot_targets: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  fs_targets_kernel<<<grid_for(static_cast<long long>(n) * num_classes, 256), 256, 0, st>>>(y, n, num_classes, out);
  FS_LAUNCH_CHECK();
  return UMD_OK;
}
extern "C" int umd_fewshot_matmul(const float* A, long long a_row_stride, long long a_col_stride, const float* B,
                                  long long b_row_stride, long long b_col_stride, float* C, long long ldc, int M, int N,
                                  int K, umd_stream_t stream) {
  UMD_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && ldc >= N, "umd_fewshot_matmul: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int gx = static_cast<int>(ceil_div_ll(N, FS_TN)), gy = static_cast<int>(ceil_div_ll(M, FS_TM));
  // split K until the grid fills the machine (Gram products: few tiles, K = support size)
  int splits = 1;
  const int tiles = gx * gy, sms = sm_count();
  if (tiles < 2 * sms && K > 4 * FS_TK * 8) {
    splits = (2 * sms) / tiles;   // one resident wave (2 CTAs per SM at 116 registers): 343 CTAs on 296 slots cost a 16 %-full second wave
    const int max_splits = K / (FS_TK * 8);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int kps = static_cast<int>(ceil_div_ll(ceil_div_ll(K, splits), FS_TK)) * FS_TK;
  splits = static_cast<int>(ceil_div_ll(K, kps));
  if (splits > 1) UMD_CHECK_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, st));
  fs_matmul_kernel<<<dim3(gx, gy, splits), 256, 0, st>>>(A, a_row_stride, a_col_stride, B, b_row_stride, b_col_stride, C,
                                                         ldc, M, N, K, kps, splits > 1 ? 1 : 0);
  FS_LAUNCH_CHECK();
  return UMD_OK;
}
extern "C" size_t umd_fewshot_solve_scratch_bytes(int n, int num_rhs) {
  if (n <= 0 || num_rhs <= 0) return 0;
  (void)num_rhs;
  return sizeof(double) * (static_cast<size_t>(n) * n) + 256;
}
extern "C" int umd_fewshot_ridge_solve(const float* gram, float l2_reg, const float* rhs, int n, int num_rhs, float* z,
                                       void* scratch, size_t scratch_bytes, int* status, umd_stream_t stream) {
  UMD_REQUIRE(gram && rhs && z && scratch && status && n > 0 && num_rhs > 0, "umd_fewshot_ridge_solve: bad argument");
  UMD_REQUIRE(n <= 2048, "umd_fewshot_ridge_solve: n = %d exceeds the shared-memory solve limit (2048)", n);
  UMD_REQUIRE(scratch_bytes >= umd_fewshot_solve_scratch_bytes(n, num_rhs), "umd_fewshot_ridge_solve: scratch too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* L = static_cast<double*>(scratch);
  UMD_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
  fs_load_system_kernel<<<grid_for(static_cast<long long>(n) * n, 256), 256, 0, st>>>(gram, l2_reg, n, L);
  FS_LAUNCH_CHECK();
  for (int kb = 0; kb < n; kb += FS_NB) {
    fs_chol_diag_kernel<<<1, 1024, 0, st>>>(L, n, kb, status);
    FS_LAUNCH_CHECK();
    const int below = n - kb - FS_NB;
    if (below <= 0) break;
    fs_chol_panel_kernel<<<static_cast<int>(ceil_div_ll(below, 128)), 128, 0, st>>>(L, n, kb);
    FS_LAUNCH_CHECK();
    const int nblk = static_cast<int>(ceil_div_ll(below, FS_NB));
    fs_chol_update_kernel<<<dim3(nblk, nblk), 1024, 0, st>>>(L, n, kb);
    FS_LAUNCH_CHECK();
  }
  fs_mirror_kernel<<<grid_for(static_cast<long long>(n) * n, 256), 256, 0, st>>>(L, n);
  FS_LAUNCH_CHECK();
  int cols = 32;
  auto smem_for = [&](int cc) { return (static_cast<size_t>(n) * (cc + FS_RING) + 128) * sizeof(double); };
  while (cols > 8 && smem_for(cols) > 227 * 1024) cols >>= 1;
  const size_t smem = smem_for(cols);
  static bool attr_set = false;
  if (!attr_set) {
    UMD_CHECK_CUDA(cudaFuncSetAttribute(fs_trisolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  UMD_REQUIRE(smem <= 227 * 1024, "umd_fewshot_ridge_solve: n = %d does not fit the shared-memory solve", n);
  fs_trisolve_kernel<<<static_cast<int>(ceil_div_ll(num_rhs, cols)), 128, smem, st>>>(L, n, num_rhs, cols, rhs, z);
  FS_LAUNCH_CHECK();
  return UMD_OK;
}
extern "C" int umd_fewshot_accuracy(const float* scores, const int* labels, int n, int num_classes, int* preds_or_null,
                                    int* correct, umd_stream_t stream) {
  UMD_REQUIRE(scores && labels && correct && n > 0 && num_classes > 0, "umd_fewshot_accuracy: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UMD_CHECK_CUDA(cudaMemsetAsync(correct, 0, sizeof(int), st));
  fs_accuracy_kernel<<<static_cast<int>(ceil_div_ll(static_cast<long long>(n) * 32, 256)), 256, 0, st>>>(
      scores, labels, n, num_classes, preds_or_null, correct);
  FS_LAUNCH_CHECK();
  return UMD_OK;
}

// =========================================================================================
// Training input stage on the GPU: the part of the reference's tf.data preprocessing string
//   decode_jpeg_and_inception_crop(size)|flip_lr|value_range(-1, 1)        (configs/ae_i1k.py:64-69)
// that follows JPEG decoding — crop window -> tf.image.resize(bilinear, antialias=False) -> clip + cast back to uint8
// (pp/ops_image.py:75-85) -> horizontal flip (:306-314) -> (x - in_min) / (in_max - in_min) * (vmax - vmin) + vmin
// (pp/ops_general.py:51-60; vmin / vmax are Python scalars there, so their difference is formed in double precision and
// rounded once: hence the double arguments) — fused into one pass: one thread per output value, uint8 in, fp32 out.
// The arithmetic follows TensorFlow's half-pixel-centre bilinear kernel operation by operation with non-contracting
// intrinsics, so the intermediate uint8 image (a truncating cast) and the fp32 result are bit-identical to an IEEE
// single-precision evaluation in the same order.  HBM-bound: <= 1 B read + 4 B written per output value.
// =========================================================================================
namespace umd {
__global__ void augment_kernel(const unsigned char* __restrict__ src, int Hs, int Ws, int C, const int* __restrict__ boxes,
                               const unsigned char* __restrict__ flips, int Sh, int Sw, float in_min, float in_max, float vmin,
                               float vmax, float span, int clip_values, float* __restrict__ out, unsigned char* __restrict__ out_u8,
                               long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long r = i / C;
    int ox = static_cast<int>(r % Sw);
    r /= Sw;
    const int oy = static_cast<int>(r % Sh);
    const int n = static_cast<int>(r / Sh);
    int y0 = 0, x0 = 0, bh = Hs, bw = Ws;
    if (boxes) { y0 = boxes[4 * n]; x0 = boxes[4 * n + 1]; bh = boxes[4 * n + 2]; bw = boxes[4 * n + 3]; }
    if (flips && flips[n]) ox = Sw - 1 - ox;
    const float sy = static_cast<float>(bh) / static_cast<float>(Sh), sx = static_cast<float>(bw) / static_cast<float>(Sw);
    const float fy = __fsub_rn(__fmul_rn(__fadd_rn(static_cast<float>(oy), 0.5f), sy), 0.5f);
    const float fx = __fsub_rn(__fmul_rn(__fadd_rn(static_cast<float>(ox), 0.5f), sx), 0.5f);
    const float fy0 = floorf(fy), fx0 = floorf(fx);
    const int yl = max(static_cast<int>(fy0), 0), yu = min(static_cast<int>(ceilf(fy)), bh - 1);
    const int xl = max(static_cast<int>(fx0), 0), xu = min(static_cast<int>(ceilf(fx)), bw - 1);
    const float ly = __fsub_rn(fy, fy0), lx = __fsub_rn(fx, fx0);
    const unsigned char* img = src + static_cast<long long>(n) * Hs * Ws * C;
    auto at = [&](int y, int x) { return static_cast<float>(img[(static_cast<long long>(y0 + y) * Ws + (x0 + x)) * C + c]); };
    const float tl = at(yl, xl), tr = at(yl, xu), bl = at(yu, xl), br = at(yu, xu);
    const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
    const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
    float v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
    v = fminf(fmaxf(v, 0.f), 255.f);
    const unsigned char q = static_cast<unsigned char>(v);         // truncation, like tf.cast(float -> uint8)
    if (out_u8) out_u8[i] = q;
    float f = __fdiv_rn(__fsub_rn(static_cast<float>(q), in_min), __fsub_rn(in_max, in_min));
    f = __fadd_rn(vmin, __fmul_rn(f, span));
    if (clip_values) f = fminf(fmaxf(f, vmin), vmax);
    out[i] = f;
  }
}
}  // namespace umd

extern "C" int umd_augment_u8(const unsigned char* images, int n, int src_h, int src_w, int channels, const int* boxes_or_null,
                              const unsigned char* flips_or_null, int out_h, int out_w, float in_min, float in_max, double vmin,
                              double vmax, int clip_values, float* out, unsigned char* resized_u8_or_null, umd_stream_t stream) {
  UMD_REQUIRE(images && out && n >= 0 && src_h > 0 && src_w > 0 && channels > 0 && out_h > 0 && out_w > 0, "umd_augment_u8: bad argument");
  UMD_REQUIRE(in_max != in_min, "umd_augment_u8: in_max == in_min");
  if (n == 0) return UMD_OK;
  const long long total = static_cast<long long>(n) * out_h * out_w * channels;
  augment_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      images, src_h, src_w, channels, boxes_or_null, flips_or_null, out_h, out_w, in_min, in_max, static_cast<float>(vmin),
      static_cast<float>(vmax), static_cast<float>(vmax - vmin), clip_values, out, resized_u8_or_null, total);
  FS_LAUNCH_CHECK();
  return UMD_OK;
}
