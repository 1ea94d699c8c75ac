/*
 * umd_b200.h — C ABI of the B200-native UMD auto-encoder training-step library
 * (libumd_b200.so).
 *
 * The reference (philippe-eecs/small-vision) is pure Python/JAX and has no FFI; the seams
 * this library replaces are Python callables:
 *   - big_vision/models/ae.py:176-197      _ViTAE.__call__      -> umd_forward
 *   - big_vision/trainers/train_ae.py:287-382  update_fn         -> umd_train_step (= umd_qsample, umd_mask_argsort,
 *                                                  umd_forward, umd_backward, umd_comm all-reduce, umd_adamw_step)
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

/* Optional live per-category device timing (bench.py roofline): while enabled the launchers bracket their
 * kernels with CUDA events on the launching stream and account algorithmic FLOPs (tensor-bound categories)
 * or bytes (HBM-bound ones).  umd_profile_read sums and clears; call it after a device synchronise.
 * Returns the number of scopes that could not be read. */
void umd_profile_enable(int on);
int umd_profile_num_categories(void);
const char* umd_profile_category_name(int category);
int umd_profile_read(float* ms, double* work, long long* scopes, int ncat);

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
  UMD_EPI_BF16_DELTA = 6, /* out0(bf16) = acc ; out1(f32)[row, c / 64] = sum over the 64-column group c of acc * aux(bf16):
                             dO = dA Wo^T together with delta = rowsum(dO o O) per head; N % 64 == 0,
                             ld1 = number of 64-column groups per row of out1 */
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
  /* b_mn = 0 only: the contraction is split over B's batch axis in chunks of b_kchunk elements (multiple of
   * 64): B element (n, k) lives at B + (k / b_kchunk) * b_bs + n * ldb + (k % b_kchunk).  0 = off.  Used for
   * dY0 = [dQ|dK|dV] [Wq|Wk|Wv]^T with the three kernels stored as separate leaves. */
  int b_kchunk;
} umd_gemm_args;

int umd_gemm_bf16(const umd_gemm_args* args, umd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stand-alone pieces of the path that the reference exposes as functions of their own.
 * ------------------------------------------------------------------------------------------ */
/* gaussian_diffusion.py:85-98  q_sample: out = sqrt_ac[t[n]] * x0 + sqrt_1mac[t[n]] * noise (fp32). */
int umd_qsample(const float* x0, const float* noise, const int* t, const float* sqrt_alphas_cumprod,
                const float* sqrt_one_minus_alphas_cumprod, int n, int elems_per_sample, float* out,
                umd_stream_t stream);
/* gaussian_diffusion.py:134-211  one DDIM update (ddim_sample over p_mean_variance), fused with the classifier-free
 * guidance combine of a doubled batch (models/ae.py:192-195) and the eps / x0 head selection of
 * trainers/train_ae.py:472-483.  x, noise, sample, pred_xstart: [n, hw, channels] fp32; pred: [n, hw, pred_channels]
 * with pred_channels = 2*channels (x0 head | eps head) or = channels (eps head only; needs eps_pred), [2n, ...] with
 * use_cfg (conditional rows first).  t_next may be null (alphas_cumprod_prev[t] is used, :187-190);
 * pred_xstart may be null.  sample = sqrt(abar_prev) x0_hat + sqrt(1 - abar_prev - sigma^2) eps + [t > 0] sigma noise. */
int umd_ddim_step(const float* x, const float* pred, const float* noise, const int* t, const int* t_next,
                  const float* alphas_cumprod, const float* alphas_cumprod_prev, const float* sqrt_recip_alphas_cumprod,
                  const float* sqrt_recipm1_alphas_cumprod, int n, int hw, int channels, int pred_channels, float eta, int use_cfg,
                  float cfg_scale, int eps_pred, int clip_denoised, float* sample, float* pred_xstart,
                  umd_stream_t stream);
/* models/ae.py:14-16,25-27  stable argsort of the mask noise, its inverse and the 0/1 sequence mask. */
int umd_mask_argsort(const float* noise, int n, int L, int len_keep, int* ids_shuffle, int* ids_restore,
                     float* mask_or_null, umd_stream_t stream);
/* models/vit.py:82-87 attention over a packed qkv buffer [rows, 3*H*Dh] (two-segment ragged batch). */
int umd_attention_fwd(const void* qkv_bf16, void* out_bf16, float* lse, int n0, int s0, int n1, int s1, int H, int Dh,
                      umd_stream_t stream);
int umd_attention_bwd(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse,
                      void* dqkv_bf16, int n0, int s0, int n1, int s1, int H, int Dh, umd_stream_t stream);
/* Backward with delta[rows, H] = rowsum(dO o O) per head supplied (the step engine gets it from the epilogue of the
 * out-projection dgrad GEMM, UMD_EPI_BF16_DELTA) instead of the forward output: 20 % fewer bytes, no prologue. */
int umd_attention_bwd_delta(const void* qkv_bf16, const void* dout_bf16, const float* lse, const float* delta,
                            void* dqkv_bf16, int n0, int s0, int n1, int s1, int H, int Dh, umd_stream_t stream);
/* same contract on the CUDA-core kernels (the in-library checker of the tcgen05 attention path) */
int umd_attention_fwd_simt(const void* qkv_bf16, void* out_bf16, float* lse, int n0, int s0, int n1, int s1, int H, int Dh,
                           umd_stream_t stream);
int umd_attention_bwd_simt(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse,
                           void* dqkv_bf16, int n0, int s0, int n1, int s1, int H, int Dh, umd_stream_t stream);
/* models/vit.py:78-80 LayerNorm (+ adaLN modulate); out bf16 when out_is_bf16 else fp32 */
int umd_ln_modulate_fwd(const float* x, const float* gamma, const float* beta, const float* shift, const float* scale,
                        long long ldmod, int n0, int s0, int n1, int s1, int D, void* out, int out_is_bf16, float* mean,
                        float* rstd, umd_stream_t stream);
/* backward: dx is written (accumulate = 0) or += (accumulate = 1); dshift, dscale (per sample), dgamma, dbeta are
 * accumulated with atomic adds — zero them first */
int umd_ln_modulate_bwd(const void* dy, int dy_is_bf16, const float* x, const float* mean, const float* rstd,
                        const float* gamma, const float* beta, const float* scale, long long ldmod, int n0, int s0,
                        int n1, int s1, int D, float* dx, int accumulate, float* dshift, float* dscale, long long ldd,
                        float* dgamma, float* dbeta, umd_stream_t stream);
/* The same backward fused with the gate backward of the branch whose residual add produced x (models/vit.py:89-94,
 * 106-108: x = x_prev + gate[sample] * z): with dx_new the value written to dx,
 *   dz = gate * dx_new (bf16, required), dgate[sample] += sum_t z * dx_new (needs z; may be null),
 *   dbias += gate * sum_t dx_new (bias of the Dense that produced z; may be null).  gate null = 1.
 * dgate / dbias are accumulated with atomic adds — zero them first. */
int umd_ln_modulate_bwd_gated(const void* dy, int dy_is_bf16, const float* x, const float* mean, const float* rstd,
                              const float* gamma, const float* beta, const float* scale, long long ldmod, int n0, int s0,
                              int n1, int s1, int D, float* dx, int accumulate, float* dshift, float* dscale, long long ldd,
                              float* dgamma, float* dbeta, void* dz_bf16, const void* z_bf16, const float* gate,
                              long long ldgate, float* dgate, long long lddgate, float* dbias, umd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser over the flat arena (train_ae.py:124-152,365-374): global-norm clip + AdamW (bf16 mu)
 * + masked weight decay + optional EMA; also refreshes the bf16 shadow of the parameters.
 * ------------------------------------------------------------------------------------------ */
typedef struct umd_adamw_args {
  float* params;            /* [n] fp32 master */
  const float* grads;       /* [n] fp32 (already all-reduced) */
  void* mu;                 /* [n] bf16 */
  float* nu;                /* [n] fp32 */
  void* params_bf16;        /* [n] bf16 shadow or NULL */
  float* ema;               /* [n] fp32 or NULL */
  const uint8_t* wd_flags;  /* [n/64] 1 = decayed leaf */
  long long n;              /* multiple of 64 */
  float clip_norm, lr, b1, b2, eps, wd;
  float bias_corr1, bias_corr2; /* 1 - b^count_inc */
  float ema_decay;
  float* scratch;           /* >= 3080 floats */
  int scratch_floats;
  float* measurements;      /* device float[3]: l2_params, l2_updates, grad_norm */
} umd_adamw_args;
int umd_adamw_step(const umd_adamw_args* args, umd_stream_t stream);
int umd_sumsq(const float* x, long long n, float* scratch, int scratch_floats, float* out, umd_stream_t stream);
/* bf16 shadow of the fp32 arena (n multiple of 4) */
int umd_cast_f32_to_bf16(const float* x, long long n, void* out_bf16, umd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The model: big_vision/models/ae.py _ViTAE (embed -> encode -> decode) and the loss of
 * trainers/train_ae.py:323-361, forward and backward, over a flat parameter arena.
 * ------------------------------------------------------------------------------------------ */
typedef struct umd_model_cfg {
  int img_size, patch, channels;
  int width, depth, dec_depth, heads, mlp_dim;
  int num_cls;
  int num_classes;      /* 0 = no label path */
  int adaln;
  int flip_final_conv;  /* 1 = flax ConvTranspose(transpose_kernel=False) orientation */
  int residual_bf16;    /* training only: 1 = the residual stream between the blocks (and its per-layer snapshots for the
                           backward) is bf16, as in the reference's dtype_mm="bfloat16" flow (ae.py:51,100; vit.py:89-94);
                           2 = so is the gradient of the residual stream between the blocks (JAX cotangents take the dtype
                           of their primals).  LayerNorm statistics, loss and optimiser stay fp32.  0 = fp32 streams. */
} umd_model_cfg;

/* Leaves of the parameter tree.  offsets[] (UMD_OFFSETS_LEN entries) are element offsets into the
 * arena (params / grads / bf16 shadow / mu / nu share one layout); -1 marks an absent leaf.
 * The scanned blocks are stored LAYER-MAJOR: a block leaf's offset is that of its layer-0 slice, layer l's slice starts
 * offsets[UMD_P_ENC_LAYER_STRIDE] (resp. _DEC_) elements further per layer, i.e. all leaves of one layer are contiguous.
 * (Flax stacks every scanned leaf over depth, vit.py:131-148; with that layout no encoder gradient would be final before
 * layer 0's backward has run.  The host-side tree presents the same [depth, ...] shapes as strided views.) */
enum {
  UMD_P_CLS = 0, UMD_P_POS, UMD_P_DEC_POS, UMD_P_MASK_TOKEN, UMD_P_EMBED_W, UMD_P_EMBED_B,
  UMD_P_TT_W0, UMD_P_TT_B0, UMD_P_TT_W1, UMD_P_TT_B1,
  UMD_P_LABEL_TABLE, UMD_P_LT_W0, UMD_P_LT_B0, UMD_P_LT_W1, UMD_P_LT_B1,
  UMD_P_FMOD_W, UMD_P_FMOD_B, UMD_P_FCONV_W, UMD_P_FCONV_B,
  UMD_P_ENC_BASE,                       /* + UMD_S_* */
  UMD_P_DEC_BASE = UMD_P_ENC_BASE + 20, /* + UMD_S_* */
  UMD_P_COUNT = UMD_P_DEC_BASE + 20,
  UMD_P_ENC_LAYER_STRIDE = UMD_P_COUNT, /* elements between consecutive encoder layer blocks */
  UMD_P_DEC_LAYER_STRIDE,               /* ... decoder layer blocks */
  UMD_OFFSETS_LEN
};
enum {
  UMD_S_ADA_W = 0, UMD_S_ADA_B, UMD_S_LN0_S, UMD_S_LN0_B, UMD_S_LN1_S, UMD_S_LN1_B,
  UMD_S_Q_W, UMD_S_K_W, UMD_S_V_W, UMD_S_Q_B, UMD_S_K_B, UMD_S_V_B, UMD_S_O_W, UMD_S_O_B,
  UMD_S_FC1_W, UMD_S_FC1_B, UMD_S_FC2_W, UMD_S_FC2_B, UMD_S_NORM_S, UMD_S_NORM_B
};

typedef struct umd_step_shape {
  int n0, n1;           /* samples in the noise segment / the clean (MAE) segment */
  int keep0, keep1;     /* kept patches per sample (L when the segment is not masked) */
  int masked0, masked1; /* 1 = random masking applies to the segment */
} umd_step_shape;

typedef struct umd_io {
  const float* image;       /* [n, H, W, C] model input (x_t for segment 0, x_0 for segment 1) */
  const int* t;             /* [n] timestep as the model sees it (t+1, or 0) */
  const int* labels;        /* [n] class ids after label-dropout / null substitution, or NULL */
  const int* ids_shuffle;   /* [n, L] */
  const int* ids_restore;   /* [n, L] */
  const float* x0;          /* [n, H, W, C] loss target, or NULL (no loss) */
  const float* noise;       /* [n0, H, W, C] loss target of the eps half */
  float* pred;              /* [n, H, W, 2C] or NULL */
  float* pre_logits;        /* [n, D] or NULL */
  float* loss;              /* device scalar or NULL */
} umd_io;

typedef void (*umd_bucket_cb)(void* user, int bucket);

size_t umd_workspace_bytes(const umd_model_cfg* cfg, const umd_step_shape* shape, int train);
/* Forward.  train != 0 keeps every activation the backward needs in the workspace. */
int umd_forward(const umd_model_cfg* cfg, const umd_step_shape* shape, const long long* offsets, const float* params,
                const void* params_bf16, const umd_io* io, void* workspace, size_t workspace_bytes, int train,
                umd_stream_t stream);
/* Backward of the loss computed by the preceding umd_forward(train=1) on the same workspace; accumulates
 * into grads (which the caller zeroes).  cb(user, event) is called on the host as soon as every kernel writing a part
 * of the gradient arena has been enqueued on `stream`: event 0 = the decoder side (final_conv, final_modulation,
 * Decoder, dec_pos_embedding, image_mask_embedding); event 1 + j, j = 0 .. depth-1 = encoder layer depth-1-j (its
 * layer block of the arena; j = 0 also covers Encoder/encoder_norm); event 1 + depth = everything (embeddings and the
 * conditioning trunks).  A data-parallel caller all-reduces the finished arena ranges behind these events while the
 * rest of the backward runs (train_ae.py:287-290,364). */
int umd_backward(const umd_model_cfg* cfg, const umd_step_shape* shape, const long long* offsets, const float* params,
                 const void* params_bf16, float* grads, const umd_io* io, void* workspace, size_t workspace_bytes,
                 umd_bucket_cb cb, void* cb_user, umd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel communicator (big_vision/sharding.py:33-55 + the implicit GSPMD gradient all-reduce of
 * trainers/train_ae.py:159-170,287-290,364).  NCCL is resolved at run time (libnccl.so.2); every collective runs on a
 * stream owned by the communicator and is ordered against the caller's streams with events.  One communicator per
 * process / GPU.  id128 is an ncclUniqueId obtained on one rank and distributed by the host (any side channel).
 * ------------------------------------------------------------------------------------------ */
int umd_comm_available(void);      /* 1 when libnccl.so.2 could be loaded */
int umd_comm_nccl_version(void);
int umd_comm_unique_id(unsigned char* id128);
int umd_comm_init(int rank, int world, const unsigned char* id128, void** comm);
/* wraps an ncclComm_t the host already owns (not destroyed by umd_comm_destroy) */
int umd_comm_from_nccl(void* nccl_comm, void** comm);
int umd_comm_world(void* comm);
int umd_comm_rank(void* comm);
/* buf[0, n) <- mean over ranks, ordered behind `stream`; `stream` waits for the result */
int umd_comm_allreduce_mean(void* comm, float* buf, long long n, umd_stream_t stream);
int umd_comm_destroy(void* comm);

/* ------------------------------------------------------------------------------------------
 * One whole training step — update_fn of big_vision/trainers/train_ae.py:287-382 — in one call: q_sample noising
 * (:318-321), masking index work (ae.py:9-28), forward of both branches, the loss (:323-361), backward (:364), the
 * data-parallel gradient mean (each bucket handed to `comm` behind the backward event that completes it), global-norm
 * clip + AdamW + EMA (:365-374) and the step measurements (:367-371).  Every random draw of the reference is an
 * argument (:302-317, ae.py:14, embeddings.py:44).  Asynchronous on `stream`; nothing is allocated.
 * ------------------------------------------------------------------------------------------ */
enum {
  UMD_STEP_NO_OPTIMIZER = 1,  /* stop after the (reduced) gradients: opt.params / mu / nu stay untouched */
};
typedef struct umd_train_step_args {
  const umd_model_cfg* cfg;
  umd_step_shape shape;            /* n0 noised + n1 clean samples; keep counts = int(L (1 - mask_ratio)) (ae.py:11) */
  const long long* offsets;        /* UMD_OFFSETS_LEN entries */
  umd_adamw_args opt;              /* params, mu, nu, params_bf16 (required), ema, hyper-parameters of this step (opt.grads is
                                      ignored); opt.measurements: device float[4] <- loss, l2_params, l2_updates, grad_norm */
  float* grads;                    /* [opt.n + 64] fp32 scratch: on return the (reduced) gradients, slot opt.n the (mean) loss */
  const float* image;              /* [B, H, W, C] x_0 in [-1, 1]; the first n0 samples take the noise branch (:307-308) */
  const long long* label;          /* [B] or NULL */
  int use_labels;
  const int* t;                    /* [n0] in [0, T) */
  const float* noise;              /* [n0, H, W, C] ~ N(0, 1) */
  const float* mask_noise0;        /* [n0, L] ~ U[0, 1), NULL when segment 0 is not masked */
  const float* mask_noise1;        /* [n1, L] */
  const unsigned char* label_drop; /* [n0] 1 = replace the label by the null class, or NULL */
  const float* sqrt_alphas_cumprod;            /* [T] */
  const float* sqrt_one_minus_alphas_cumprod;  /* [T] */
  void* workspace;                 /* umd_train_workspace_bytes(cfg, shape) bytes */
  size_t workspace_bytes;
  void* comm;                      /* NULL = single GPU */
  const long long* bucket_bounds;  /* [num_buckets][2] element ranges of grads, each reduced behind ... */
  const int* bucket_events;        /* [num_buckets] ... this backward event (see umd_backward); a range that ends at
                                      opt.n is extended by the 64 trailing slots */
  int num_buckets;
  int flags;
} umd_train_step_args;
size_t umd_train_workspace_bytes(const umd_model_cfg* cfg, const umd_step_shape* shape);
int umd_train_step(const umd_train_step_args* args, umd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Few-shot ridge probe on `pre_logits`: big_vision/evaluators/fewshot_lsr.py.
 * All buffers fp32 / int32 device pointers, row-major.  dim = d + 1 (the appended bias feature).
 * ------------------------------------------------------------------------------------------ */
/* fewshot_lsr.py:46-47  column mean and population std + 1e-5 of the support features x[n, d]. */
int umd_fewshot_stats(const float* x, int n, int d, float* mean, float* std_plus_eps, umd_stream_t stream);
/* fewshot_lsr.py:48-51 (support) and :99-100 (query): out[n, d + 1] = [(x - mean) / std | bias_constant]. */
int umd_fewshot_whiten(const float* x, const float* mean, const float* std_plus_eps, int n, int d, float bias_constant,
                       float* out, umd_stream_t stream);
/* fewshot_lsr.py:54,82  rhs[dim, C] = X^T (2 onehot(y) - 1) from per-class row sums; sums_scratch: (C + 2) * dim floats. */
int umd_fewshot_xty(const float* xw, const int* y, int n, int dim, int num_classes, float* sums_scratch, float* rhs,
                    umd_stream_t stream);
/* fewshot_lsr.py:54  out[n, C] = 2 onehot(y) - 1 (needed as the right-hand side when n < dim, :86-88). */
int umd_fewshot_targets(const int* y, int n, int num_classes, float* out, umd_stream_t stream);
/* fp32 product with element strides, C[m, n] = sum_k A[m * a_row_stride + k * a_col_stride] * B[k * b_row_stride +
 * n * b_col_stride]: x^T x / x x^T (fewshot_lsr.py:81,85), x^T z (:88,108) and x_test w (:111) are all this call. */
int umd_fewshot_matmul(const float* A, long long a_row_stride, long long a_col_stride, const float* B,
                       long long b_row_stride, long long b_col_stride, float* C, long long ldc, int M, int N, int K,
                       umd_stream_t stream);
/* fewshot_lsr.py:103-108  z = (gram + l2_reg I)^-1 rhs.  The reference applies the inverse through the cached
 * eigendecomposition (q diag(1 / (eigs + l2)) q^T); here a Cholesky factorisation in fp64 gives the same solution.
 * gram [n, n] symmetric, rhs / z [n, num_rhs] (may alias); *status (device int) = 0, or 1 + the first non-positive
 * pivot.  n <= 2048. */
size_t umd_fewshot_solve_scratch_bytes(int n, int num_rhs);
int umd_fewshot_ridge_solve(const float* gram, float l2_reg, const float* rhs, int n, int num_rhs, float* z, void* scratch,
                            size_t scratch_bytes, int* status, umd_stream_t stream);
/* fewshot_lsr.py:111-112  preds = argmax(scores[n, C], axis 1) (first maximum); *correct = #(preds == labels). */
int umd_fewshot_accuracy(const float* scores, const int* labels, int n, int num_classes, int* preds_or_null, int* correct,
                         umd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Training input stage after JPEG decoding, one fused pass over a uint8 batch:
 * crop window (pp/ops_image.py:197-242; boxes[n] = {y0, x0, h, w}, null = whole image) -> bilinear resize to out_h x out_w
 * with half-pixel centres, clipped and truncated back to uint8 (pp/ops_image.py:75-85) -> horizontal flip where
 * flips[n] != 0 (:306-314) -> value_range (pp/ops_general.py:51-60).  images [n, src_h, src_w, C] uint8;
 * out [n, out_h, out_w, C] fp32; resized_u8 (optional) receives the intermediate uint8 image.  vmin / vmax are doubles because
 * the reference forms (vmax - vmin) from Python scalars before rounding to float32.  The crop boxes and flip
 * flags are inputs (the reference draws them with tf.image.sample_distorted_bounding_box / random_flip_left_right).
 * ------------------------------------------------------------------------------------------ */
int umd_augment_u8(const unsigned char* images, int n, int src_h, int src_w, int channels, const int* boxes_or_null,
                   const unsigned char* flips_or_null, int out_h, int out_w, float in_min, float in_max, double vmin,
                   double vmax, int clip_values, float* out, unsigned char* resized_u8_or_null, umd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UMD_B200_H_ */
