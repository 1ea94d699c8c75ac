/*
 * umd_b200.h — C ABI of the B200-native UMD auto-encoder training-step library
 * (libumd_b200.so).
 *
 * The reference (philippe-eecs/small-vision) is pure Python/JAX and has no FFI; the seams
 * this library replaces are Python callables:
 *   - big_vision/models/ae.py:176-197      _ViTAE.__call__      -> umd_forward
 *   - big_vision/trainers/train_ae.py:287-382  update_fn         -> umd_train_step + umd_adamw_step
 *   - big_vision/gaussian_diffusion.py:85-98   q_sample          -> umd_qsample
 *   - big_vision/models/ae.py:9-28         random_masking        -> umd_mask_argsort
 * Every function is asynchronous on the caller's cudaStream_t, allocates nothing, takes raw
 * device pointers and plain sizes, returns 0 on success, and leaves a message retrievable by
 * umd_last_error() otherwise.  A JAX binding wraps these with XLA_FFI_DEFINE_HANDLER (see
 * INTEGRATION.md); in this repo they are driven from Python through ctypes.
 */
#ifndef UMD_B200_H_
#define UMD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* umd_stream_t; /* cudaStream_t */

enum {
  UMD_OK = 0,
  UMD_ERR_INVALID = 1,
  UMD_ERR_CUDA = 2,
  UMD_ERR_UNSUPPORTED = 3,
};

const char* umd_last_error(void);
int umd_version(void);
/* number of kernels this library has launched since load (all streams); bench.py reads it */
long long umd_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Dense bf16 GEMM on tcgen05/TMEM fed by TMA (models/vit.py:54,57,71,82-87; ae.py:64,94,95).
 *   D[b] = op(A[b]) * op(B[b])   (fp32 accumulate), b = 0..batch-1
 * A is M x K.  a_mn = 0: stored [M][lda] (K contiguous);  a_mn = 1: stored [K][lda] (M contiguous).
 * B is K x N.  b_mn = 0: stored [N][ldb] (K contiguous);  b_mn = 1: stored [K][ldb] (N contiguous).
 * Batch strides are in elements; 0 broadcasts the operand.
 * ------------------------------------------------------------------------------------------ */
enum {
  UMD_EPI_BF16 = 0,     /* out0(bf16) = acc + bias                                             */
  UMD_EPI_F32 = 1,      /* out0(f32)  = acc + bias                                             */
  UMD_EPI_GELU = 2,     /* out0(bf16) = u = acc + bias ; out1(bf16) = gelu_tanh(u)   vit.py:54-55 */
  UMD_EPI_GATE_RES = 3, /* out0(bf16, optional) = a = acc + bias ;
                           out1(f32) = aux(f32) + gate[sample(row)] * a        vit.py:89-94,106-108 */
  UMD_EPI_DGELU = 4,    /* out0(bf16) = acc * gelu_tanh'(aux(bf16))                             */
  UMD_EPI_ATOMIC = 5,   /* out0(f32) += acc   (split-K weight gradients)                        */
};

typedef struct umd_gemm_args {
  const void* A;
  const void* B;
  int M, N, K, batch;
  int a_mn, b_mn;
  long long lda, ldb;
  long long a_bs, b_bs;
  int epi;
  int split_k;        /* >1 only with UMD_EPI_ATOMIC */
  void* out0;
  long long ld0, bs0;
  void* out1;
  long long ld1;
  const float* bias;  /* [N] or NULL */
  long long bias_bs;
  const void* aux;
  long long ldaux;
  const float* gate;  /* per-sample rows of length >= N, or NULL (= 1) */
  long long ldgate;
  /* row -> sample map for the gate (two-segment ragged batch) */
  int split_row, s0, s1, n0;
} umd_gemm_args;

int umd_gemm_bf16(const umd_gemm_args* args, umd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UMD_B200_H_ */
