// Internal C++ interface between the kernel translation units and the step engine.
#pragma once
#include "common.cuh"

namespace umd {

// ---- GEMM (gemm.cu) -------------------------------------------------------------------
int gemm_bf16(const umd_gemm_args& a, cudaStream_t stream);

// ---- LayerNorm + modulate (elementwise.cu) --------------------------------------------
struct LnFwdArgs {
  const void* x;           // [rows_in, D] residual stream: fp32, or bf16 when x_bf16 (umd_model_cfg.residual_bf16)
  int x_bf16;
  int xout_bf16;           // dtype of x_out
  const float* gamma;      // [D]
  const float* beta;       // [D]
  const float* shift;      // per-sample rows (stride ldmod) or null
  const float* scale;      // per-sample rows (stride ldmod) or null
  long long ldmod;
  RowMap rm;               // row -> sample map of x
  void* out;               // [rows_out, D] bf16 or fp32
  float* mean;             // [rows_out] or null
  float* rstd;             // [rows_out] or null
  int rows_out;
  int gather_L;            // >0: out row (n, j<gather_L) reads x row (n, gather_off + j)
  int gather_off;
  // Optional fused residual update (vit.py:89-94,106-108) in front of the LayerNorm:
  //   x_new = x + res_gate[sample] * res_branch   (res_gate null = 1),   written to x_out (may alias x)
  const __nv_bfloat16* res_branch;  // [rows_in, D] or null
  const float* res_gate;            // per-sample rows (stride ldgate) or null
  long long ldgate;
  void* x_out;                      // [rows_in, D] (fp32 / bf16, see xout_bf16) or null
  // adaln=False (vit.py:73-74): token 0 of every sample is replaced by cond_row[sample] before the norm
  const float* cond_row;            // [nsamples, D] or null
};
int ln_mod_fwd(const LnFwdArgs& a, int D, bool out_bf16, cudaStream_t st);

struct LnBwdArgs {
  const void* dy;          // [rows_out, D] bf16 or fp32 (gradient of the LN(+mod) output)
  const void* x;           // LN input: fp32, or bf16 when x_bf16
  int x_bf16;
  const float* mean;
  const float* rstd;
  const float* gamma;
  const float* beta;
  const float* scale;      // per-sample (1+scale) factor or null
  long long ldmod;
  RowMap rm;
  void* dx;                // [rows_in, D] gradient of the residual stream (fp32, or bf16 when dxout_bf16): written
  const void* dx_in;       // with accumulate: the running gradient that is added (may be dx itself: in place); dxin_bf16
  int dxin_bf16, dxout_bf16;
  int accumulate;
  float* dshift;           // per-sample outputs (stride ldd) or null
  float* dscale;
  long long ldd;
  float* dgamma;           // [D] atomically accumulated
  float* dbeta;
  int gather_L;
  int gather_off;
  // adaln=False (vit.py:73-74,111-112): the gradient of every sample's token-0 row goes to dcond[n] and the
  // row itself is cleared (the block's output row 0 was discarded)
  float* dcond;
  // Optional fused gate backward (App. E steps 1 and 6) of the branch whose residual add produced x:
  //   dz = gate * dx_new (bf16 GEMM operand), dgate[n] = sum_t z dx_new, dbias += gate * sum_t dx_new
  __nv_bfloat16* g_dz;        // [rows, D] out; null = no gate stage
  const __nv_bfloat16* g_z;   // [rows, D] saved branch output (needed only with g_dgate)
  const float* g_gate;        // per-sample rows (stride g_ldgate) or null (= 1)
  long long g_ldgate;
  float* g_dgate;             // per-sample out (stride g_lddgate) or null
  long long g_lddgate;
  float* g_dbias;             // [D] atomically accumulated or null
};
int ln_mod_bwd(const LnBwdArgs& a, int D, int nsamples, bool dy_bf16, cudaStream_t st);

struct GateBwdArgs {
  const float* dx;         // [rows, D] fp32 gradient of the residual stream
  const __nv_bfloat16* z;  // [rows, D] saved branch output (A or Z)
  const float* gate;       // per-sample gate rows (stride ldgate) or null (=1)
  long long ldgate;
  RowMap rm;
  __nv_bfloat16* dz;       // [rows, D] bf16 out: gate * dx
  float* dgate;            // per-sample out (stride lddgate) or null
  long long lddgate;
  float* dbias;            // [D] atomically accumulated (bias of the producing Dense) or null
};
int gate_bwd(const GateBwdArgs& a, int D, int nsamples, cudaStream_t st);

// seg_cols > 0: column c is accumulated into out[(c / seg_cols) * seg_stride + c % seg_cols]
int colsum_bf16(const void* x, long long ld, int rows, int N, float* out, cudaStream_t st, int seg_cols = 0,
                long long seg_stride = 0);
int colsum_f32(const float* x, long long ld, int rows, int N, float* out, cudaStream_t st);

// ---- conditioning path ----------------------------------------------------------------
int time_embed(const int* t, int B, int D, void* out_bf16, cudaStream_t st);
int silu_cast(const float* h, long long n, void* out_bf16, cudaStream_t st);
int silu_bwd(const void* da_bf16, const float* h, long long n, void* dh_bf16, cudaStream_t st);
int cond_combine(const float* tc, const float* yc, long long n, int adaln, float* s_out, float* cond_f32,
                 void* cond_bf16, cudaStream_t st);
int cond_combine_bwd(const float* dcond, const float* s, long long n, int adaln, void* ds_bf16, cudaStream_t st);
int gather_rows(const float* table, const int* ids, int B, int D, void* out_bf16, cudaStream_t st);
int scatter_add_rows(const float* d, const int* ids, int B, int D, float* dtable, cudaStream_t st);
int cast_bf16(const float* x, long long n, void* out, cudaStream_t st);

// ---- data movement --------------------------------------------------------------------
int qsample(const float* x0, const float* noise, const int* t, const float* sqrt_ac, const float* sqrt_1mac, int n,
            int per_sample, float* out, cudaStream_t st);
int mask_argsort(const float* noise, int n, int L, int keep, int* ids_shuffle, int* ids_restore, float* mask,
                 cudaStream_t st);
struct DdimArgs {
  const float* x;        // [n, hw, C] current sample x_t
  const float* pred;     // [n or 2n, hw, 2C] model output (x0 head, eps head)
  const float* noise;    // [n, hw, C] N(0,1) draws
  const int* t;          // [n] current timestep
  const int* t_next;     // [n] next timestep, or null (alphas_cumprod_prev[t] is used)
  const float* ac;       // alphas_cumprod
  const float* ac_prev;  // alphas_cumprod_prev
  const float* sqrt_recip_ac;
  const float* sqrt_recipm1_ac;
  int n, hw, C;
  int pred_ld;           // values per pixel in pred: 2C (x0 head, eps head) or C (eps head only)
  float eta, cfg_scale;
  int use_cfg, eps_pred, clip_denoised;
  float* sample;         // [n, hw, C]
  float* pred_xstart;    // [n, hw, C] or null
};
int ddim_step(const DdimArgs& a, cudaStream_t st);

struct EmbedArgs {
  const float* image;      // [n, img, img, C] fp32 NHWC
  const int* ids_keep;     // [n, L] ids_shuffle (first keep entries used) or null if nothing is masked
  const float* W;          // [p*p*C, D] fp32 (Flax conv kernel [p,p,C,D])
  const float* bias;       // [D]
  const float* pos;        // [L, D]
  const float* cls;        // [num_cls, D]
  float* x;                // encoder residual stream [rows, D]
  RowMap rm;
  int n1;
  int keep0, keep1, masked0, masked1;
  int img, patch, C, D, L, num_cls, tok0;
};
int embed_fwd(const EmbedArgs& a, int nsamples, cudaStream_t st);
int embed_bwd(const EmbedArgs& a, int nsamples, const float* dx, float* dW, float* db, float* dpos, float* dcls,
              cudaStream_t st);
int set_cond_row(float* x, const float* cond, const RowMap& rm, int nsamples, int D, cudaStream_t st);
int cond_row_bwd(float* dx, float* dcond, const RowMap& rm, int nsamples, int D, cudaStream_t st);

struct DecInArgs {
  const float* enc;        // encoder final-LN output [rows_enc, D] fp32
  const int* ids_restore;  // [n, L]
  const int* ids_keep;     // [n, L]
  const float* mask_token; // [D]
  const float* dec_pos;    // [L, D]
  float* xd;               // decoder residual stream [n * S_d, D]
  float* rep;              // [n, D] or null
  RowMap rm_enc;
  int keep0, keep1, masked0, masked1;
  int D, L, num_cls, tok0, S_d;
};
int decoder_input_fwd(const DecInArgs& a, int nsamples, cudaStream_t st);
int decoder_input_bwd(const DecInArgs& a, int nsamples, int enc_rows, const float* dxd, float* denc, float* ddec_pos,
                      float* dmask_token, cudaStream_t st);

struct LossArgs {
  const float* predp;      // [n*L, p*p*2C] fp32 (bias included)
  const float* x0;         // [n, img, img, C]
  const float* noise;      // [n0, img, img, C]
  const int* ids_restore;  // [n, L]
  __nv_bfloat16* dpredp;   // [n*L, p*p*2C] or null
  float* partials;         // [loss_num_partials]
  int n0, n1, keep0, keep1, masked0, masked1;
  int img, patch, C, L;
  float w_x0_0, w_eps_0, w_x0_1;
  float grad_scale;
};
int loss_num_partials(const LossArgs& a);
int loss_fwd_bwd(const LossArgs& a, float* loss_out, cudaStream_t st);
int unpatchify(const float* predp, int n, int img, int patch, int C2, float* pred, cudaStream_t st);
int pack_final_conv(const float* K, const float* bias, int p, int D, int C2, int flip, void* Wm, float* biasm,
                    cudaStream_t st);
int unpack_final_conv_grad(const float* dWm, const float* dbiasm, int p, int D, int C2, int flip, float* dK,
                           float* dbias, cudaStream_t st);

// ---- attention (attention.cu) ---------------------------------------------------------
struct AttnArgs {
  const __nv_bfloat16* qkv;   // [rows, 3*H*Dh]: q | k | v, each [H, Dh]
  __nv_bfloat16* out;         // [rows, H*Dh]
  float* lse;                 // [rows, H] log-sum-exp of the scaled logits (natural log)
  RowMap rm;
  int nsamples, H, Dh;
  float scale;                // 1/sqrt(Dh)
};
struct AttnBwdArgs {
  const __nv_bfloat16* qkv;
  const __nv_bfloat16* out;   // [rows, H*Dh] forward output, or null when delta is supplied
  const float* delta;         // [rows, H] rowsum(dO o O) per head (produced by the out-projection dgrad epilogue), or null
  const __nv_bfloat16* dout;  // [rows, H*Dh]
  const float* lse;
  __nv_bfloat16* dqkv;        // [rows, 3*H*Dh]
  RowMap rm;
  int nsamples, H, Dh;
  float scale;
};
int attention_fwd(const AttnArgs& a, cudaStream_t st);
int attention_bwd(const AttnBwdArgs& a, cudaStream_t st);
// CUDA-core kernels: in-library checker of the tensor-core path and fallback for shapes it does not cover
int attention_fwd_simt(const AttnArgs& a, cudaStream_t st);
int attention_bwd_simt(const AttnBwdArgs& a, cudaStream_t st);

// model-side timestep / label vectors of one training step (train_ae.py:327,341; ae.py:107-110; embeddings.py:43-45):
// t_model[i] = t[i] + 1 for the n0 noised samples, 0 for the clean ones; labels_model[i] = label[i] (the null class
// num_classes where label_drop[i] != 0) for the noised samples when use_labels, else the null class
int step_prep(const int* t, const long long* label, const unsigned char* label_drop, int n0, int B, int num_classes,
              int use_labels, int* t_model, int* labels_model, cudaStream_t st);

// ---- data-parallel communicator (comm.cu) -------------------------------------------------
struct Comm;
int comm_world(const Comm* c);
int comm_allreduce_mean_after(Comm* c, float* buf, long long n, cudaStream_t after, cudaEvent_t after2);
int comm_join(Comm* c, cudaStream_t stream);

// ---- optimiser (optimizer.cu) -----------------------------------------------------------
int sumsq(const float* x, long long n, float* partials, int max_partials, float* out, cudaStream_t st);

}  // namespace umd
