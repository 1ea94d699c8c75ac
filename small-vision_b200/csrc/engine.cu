// Step engine: host-side orchestration of the UMD auto-encoder forward and backward over a flat
// parameter arena and a caller-supplied workspace.  Follows big_vision/models/ae.py:99-197
// (embed -> encode -> decode), big_vision/models/vit.py:60-163 (blocks) and the loss of
// big_vision/trainers/train_ae.py:323-361; backward per SURVEY.md App. E.  Nothing here allocates or
// synchronises; every kernel goes to the caller's stream.
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace umd {

namespace {

struct Bump {
  uint8_t* base;
  size_t off;
  explicit Bump(void* b) : base(static_cast<uint8_t*>(b)), off(0) {}
  template <typename T>
  T* take(long long count) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += static_cast<size_t>(count < 0 ? 0 : count) * sizeof(T);
    return p;
  }
};

typedef __nv_bfloat16 bf16;

struct LayerBufs {
  bf16* y0; float* mean0; float* rstd0;
  bf16* qkv; bf16* o; float* lse; bf16* a;
  void* xmid;               // fp32, or bf16 with Stack::rbf
  bf16* y1; float* mean1; float* rstd1;
  bf16* u; bf16* g; bf16* z;
};

struct Stack {
  int depth, rows, nsamples, D;
  RowMap rm;
  int base;                 // leaf id base (UMD_P_ENC_BASE / UMD_P_DEC_BASE)
  long long lstride;        // arena stride between the layer blocks of this stack (layer-major parameter arena)
  float* ada;               // [depth][B][6D] fp32 or null: adaLN shift0|scale0|gate0|shift1|scale1|gate1 per layer and sample
  float* dada;              // [depth][B][6D] fp32 gradient
  bf16* dada_bf16;
  int rbf;                  // bf16 residual stream (umd_model_cfg.residual_bf16, training only): x[1..depth] and xmid are bf16
  std::vector<void*> x;     // depth+1 residual snapshots; x[0] (written by the embedding kernels) is always fp32
  std::vector<LayerBufs> L;
  float* xf; float* meanf; float* rstdf;   // final LayerNorm (encoder: fp32 output; decoder: see Plan.xm)
};

struct Plan {
  // geometry
  int B, D, H, Dh, M4, L, p, C, NC, tok0, S0, S1, Sd, Te, Td, ncls;
  int adaln, has_label;
  int rbf;   // umd_model_cfg.residual_bf16
  // conditioning path
  bf16* temb; float* th1; bf16* ta1; float* tc;
  bf16* lemb; float* lh1; bf16* la1; float* yc;
  float* s; float* cond; bf16* cond_bf16;
  float* fmod; float* dfmod; bf16* dfmod_bf16;
  Stack enc, dec;
  // decoder tail
  bf16* xm; float* meanF; float* rstdF;
  bf16* Wf_mat; float* biasm; float* predp; bf16* dpredp; float* loss_partials; int n_loss_partials;
  // backward scratch
  float* dx_dec; float* dx_enc; float* dxf_enc;
  bf16* dxb_dec; bf16* dxb_enc;   // running gradient streams in bf16 (residual_bf16 == 2), else null
  bf16* dzb; bf16* dgb; bf16* dyb; bf16* dqkv; float* delta;
  float* dcond; bf16* ds; bf16* dta1; bf16* dth1; bf16* dla1; bf16* dlh1; float* dlemb;
  float* dWf_mat; float* dbiasm;
  size_t bytes;
};

int g_sm = 0;

// ------------------------------------------------------------------------------------------
// Side stream.  The bias-gradient column sums (fc1 bias from the dGELU output, q/k/v biases from dqkv) are pure
// HBM passes over tensors that a GEMM has just written and that two more GEMMs are about to read.  They run on a
// library-owned low-priority stream, forked behind their producer and joined at the end of the layer, so that
// their CTAs (no shared memory, 34 registers) share the SMs with the tensor-bound GEMMs of the main stream instead
// of taking 2.4 ms of the step for themselves.  One caller per device (see the header), hence plain statics.
// ------------------------------------------------------------------------------------------
struct Side {
  cudaStream_t st = nullptr;
  cudaEvent_t ev[64];
  int next = 0;
  int state = -1;   // -1 unknown, 0 off, 1 ready
  int device = -1;
};
Side g_side;
int g_side_enabled = -1;   // UMD_SIDE_STREAM=0 / umd_debug_side_stream(0) keep everything on the caller's stream

bool side_ready() {
  if (g_side_enabled < 0) {
    const char* e = getenv("UMD_SIDE_STREAM");
    g_side_enabled = e ? (atoi(e) != 0) : 1;
  }
  if (!g_side_enabled) return false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (g_side.state == 1 && g_side.device == dev) return true;
  if (g_side.state == 1) return false;   // created for another device: stay on the caller's stream
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);   // lo = least priority
  if (cudaStreamCreateWithPriority(&g_side.st, cudaStreamNonBlocking, lo) != cudaSuccess) { g_side.state = 0; g_side_enabled = 0; return false; }
  for (int i = 0; i < 64; ++i)
    if (cudaEventCreateWithFlags(&g_side.ev[i], cudaEventDisableTiming) != cudaSuccess) { g_side.state = 0; g_side_enabled = 0; return false; }
  g_side.state = 1;
  g_side.device = dev;
  return true;
}
cudaEvent_t side_event() { return g_side.ev[g_side.next++ & 63]; }

// Runs fn(stream) behind everything enqueued on `main` so far, on the side stream when there is one; returns the event
// that marks its completion (null when it ran on `main`)
template <typename F>
int on_side(cudaStream_t main, cudaEvent_t* done, F&& fn) {
  *done = nullptr;
  if (!side_ready()) return fn(main);
  cudaEvent_t e = side_event();
  UMD_CHECK_CUDA(cudaEventRecord(e, main));
  UMD_CHECK_CUDA(cudaStreamWaitEvent(g_side.st, e, 0));
  UMD_TRY(fn(g_side.st));
  cudaEvent_t d = side_event();
  UMD_CHECK_CUDA(cudaEventRecord(d, g_side.st));
  *done = d;
  return UMD_OK;
}
int join_side(cudaStream_t main, cudaEvent_t done) {
  if (done) UMD_CHECK_CUDA(cudaStreamWaitEvent(main, done, 0));
  return UMD_OK;
}

void carve_stack(Bump& b, Stack& s, const Plan& P, int depth, int rows, int nsamples, const RowMap& rm, int base,
                 bool train) {
  s.depth = depth; s.rows = rows; s.nsamples = nsamples; s.rm = rm; s.base = base;
  const int D = P.D;
  s.D = D;
  s.ada = P.adaln ? b.take<float>(static_cast<long long>(P.B) * depth * 6 * D) : nullptr;
  s.dada = (P.adaln && train) ? b.take<float>(static_cast<long long>(P.B) * depth * 6 * D) : nullptr;
  s.dada_bf16 = (P.adaln && train) ? b.take<bf16>(static_cast<long long>(P.B) * depth * 6 * D) : nullptr;
  s.x.assign(depth + 1, nullptr);
  s.L.assign(depth, LayerBufs());
  const long long RD = static_cast<long long>(rows) * D;
  s.rbf = (train && P.rbf) ? 1 : 0;
  if (train) {
    s.x[0] = b.take<float>(RD);
    for (int l = 1; l <= depth; ++l) s.x[l] = s.rbf ? static_cast<void*>(b.take<bf16>(RD)) : static_cast<void*>(b.take<float>(RD));
  } else {
    float* x = b.take<float>(RD);
    for (int l = 0; l <= depth; ++l) s.x[l] = x;  // GATE_RES epilogue is element-wise in place
  }
  for (int l = 0; l < depth; ++l) {
    LayerBufs& lb = s.L[l];
    if (train || l == 0) {
      lb.y0 = b.take<bf16>(RD); lb.mean0 = b.take<float>(rows); lb.rstd0 = b.take<float>(rows);
      lb.qkv = b.take<bf16>(3 * RD); lb.o = b.take<bf16>(RD);
      lb.lse = b.take<float>(static_cast<long long>(rows) * P.H);
      lb.a = b.take<bf16>(RD);
      lb.xmid = train ? (s.rbf ? static_cast<void*>(b.take<bf16>(RD)) : static_cast<void*>(b.take<float>(RD))) : s.x[0];
      lb.y1 = b.take<bf16>(RD); lb.mean1 = b.take<float>(rows); lb.rstd1 = b.take<float>(rows);
      lb.u = b.take<bf16>(static_cast<long long>(rows) * P.M4);
      lb.g = b.take<bf16>(static_cast<long long>(rows) * P.M4);
      lb.z = b.take<bf16>(RD);
    } else {
      lb = s.L[0];
    }
  }
}

int make_plan(Plan& P, const umd_model_cfg& c, const umd_step_shape& sh, void* ws, bool train) {
  UMD_REQUIRE(c.width % 128 == 0 && c.width <= 1024, "width %d must be a multiple of 128 and <= 1024", c.width);
  UMD_REQUIRE(c.width % c.heads == 0 && c.width / c.heads == 64, "head dim must be 64 (width %d, heads %d)", c.width, c.heads);
  UMD_REQUIRE(c.img_size % c.patch == 0, "img_size %d not divisible by patch %d", c.img_size, c.patch);
  UMD_REQUIRE(c.depth >= 1 && c.dec_depth >= 1, "depth %d / dec_depth %d must be at least 1", c.depth, c.dec_depth);
  UMD_REQUIRE(sh.n0 >= 0 && sh.n1 >= 0 && sh.n0 + sh.n1 > 0, "empty batch");
  P.B = sh.n0 + sh.n1; P.D = c.width; P.H = c.heads; P.Dh = c.width / c.heads;
  P.M4 = c.mlp_dim > 0 ? c.mlp_dim : 4 * c.width;
  P.L = (c.img_size / c.patch) * (c.img_size / c.patch); P.p = c.patch; P.C = c.channels;
  P.NC = c.patch * c.patch * 2 * c.channels;
  UMD_REQUIRE(P.NC % 8 == 0, "patch*patch*2*channels = %d must be a multiple of 8", P.NC);
  UMD_REQUIRE(P.M4 % 64 == 0, "mlp_dim %d must be a multiple of 64", P.M4);
  P.adaln = c.adaln; P.has_label = c.num_classes > 0; P.ncls = c.num_cls; P.rbf = c.residual_bf16 ? 1 : 0;
  P.tok0 = c.adaln ? 0 : 1;
  UMD_REQUIRE(sh.keep0 <= P.L && sh.keep1 <= P.L && (sh.n0 == 0 || sh.keep0 > 0) && (sh.n1 == 0 || sh.keep1 > 0), "bad keep counts");
  UMD_REQUIRE(sh.masked0 || sh.n0 == 0 || sh.keep0 == P.L, "unmasked segment 0 must keep all %d patches", P.L);
  UMD_REQUIRE(sh.masked1 || sh.n1 == 0 || sh.keep1 == P.L, "unmasked segment 1 must keep all %d patches", P.L);
  P.S0 = P.tok0 + c.num_cls + sh.keep0; P.S1 = P.tok0 + c.num_cls + sh.keep1;
  P.Sd = P.tok0 + 1 + P.L;
  P.Te = sh.n0 * P.S0 + sh.n1 * P.S1; P.Td = P.B * P.Sd;
  const int B = P.B, D = P.D;
  Bump b(ws);
  P.temb = b.take<bf16>(static_cast<long long>(B) * D); P.th1 = b.take<float>(static_cast<long long>(B) * 2 * D);
  P.ta1 = b.take<bf16>(static_cast<long long>(B) * 2 * D); P.tc = b.take<float>(static_cast<long long>(B) * D);
  if (P.has_label) {
    P.lemb = b.take<bf16>(static_cast<long long>(B) * D); P.lh1 = b.take<float>(static_cast<long long>(B) * 2 * D);
    P.la1 = b.take<bf16>(static_cast<long long>(B) * 2 * D); P.yc = b.take<float>(static_cast<long long>(B) * D);
  } else {
    P.lemb = nullptr; P.lh1 = nullptr; P.la1 = nullptr; P.yc = nullptr;
  }
  P.s = b.take<float>(static_cast<long long>(B) * D); P.cond = b.take<float>(static_cast<long long>(B) * D);
  P.cond_bf16 = b.take<bf16>(static_cast<long long>(B) * D);
  P.fmod = P.adaln ? b.take<float>(static_cast<long long>(B) * 2 * D) : nullptr;
  P.dfmod = (P.adaln && train) ? b.take<float>(static_cast<long long>(B) * 2 * D) : nullptr;
  P.dfmod_bf16 = (P.adaln && train) ? b.take<bf16>(static_cast<long long>(B) * 2 * D) : nullptr;
  carve_stack(b, P.enc, P, c.depth, P.Te, B, ragged_rowmap(sh.n0, P.S0, sh.n1, P.S1), UMD_P_ENC_BASE, train);
  P.enc.xf = b.take<float>(static_cast<long long>(P.Te) * D);
  P.enc.meanf = b.take<float>(P.Te); P.enc.rstdf = b.take<float>(P.Te);
  carve_stack(b, P.dec, P, c.dec_depth, P.Td, B, uniform_rowmap(B, P.Sd), UMD_P_DEC_BASE, train);
  P.dec.xf = nullptr; P.dec.meanf = nullptr; P.dec.rstdf = nullptr;
  const long long BL = static_cast<long long>(B) * P.L;
  P.xm = b.take<bf16>(BL * D); P.meanF = b.take<float>(BL); P.rstdF = b.take<float>(BL);
  P.Wf_mat = b.take<bf16>(static_cast<long long>(D) * P.NC); P.biasm = b.take<float>(P.NC);
  P.predp = b.take<float>(BL * P.NC);
  P.dpredp = train ? b.take<bf16>(BL * P.NC) : nullptr;
  P.n_loss_partials = static_cast<int>(ceil_div_ll(BL * P.p * P.p, 256));
  P.loss_partials = b.take<float>(P.n_loss_partials);
  if (train) {
    const long long Tmax = P.Te > P.Td ? P.Te : P.Td;
    P.dx_dec = b.take<float>(static_cast<long long>(P.Td) * D);
    P.dx_enc = b.take<float>(static_cast<long long>(P.Te) * D);
    P.dxb_dec = P.rbf >= 2 ? b.take<bf16>(static_cast<long long>(P.Td) * D) : nullptr;
    P.dxb_enc = P.rbf >= 2 ? b.take<bf16>(static_cast<long long>(P.Te) * D) : nullptr;
    P.dxf_enc = b.take<float>(static_cast<long long>(P.Te) * D);
    P.dzb = b.take<bf16>(Tmax * D); P.dgb = b.take<bf16>(Tmax * P.M4); P.dyb = b.take<bf16>(Tmax * D);
    P.dqkv = b.take<bf16>(Tmax * 3 * D);
    P.delta = b.take<float>(Tmax * P.H);
    P.dcond = b.take<float>(static_cast<long long>(B) * D); P.ds = b.take<bf16>(static_cast<long long>(B) * D);
    P.dta1 = b.take<bf16>(static_cast<long long>(B) * 2 * D); P.dth1 = b.take<bf16>(static_cast<long long>(B) * 2 * D);
    P.dla1 = P.has_label ? b.take<bf16>(static_cast<long long>(B) * 2 * D) : nullptr;
    P.dlh1 = P.has_label ? b.take<bf16>(static_cast<long long>(B) * 2 * D) : nullptr;
    P.dlemb = P.has_label ? b.take<float>(static_cast<long long>(B) * D) : nullptr;
    P.dWf_mat = b.take<float>(static_cast<long long>(D) * P.NC); P.dbiasm = b.take<float>(P.NC);
  }
  P.bytes = (b.off + 255) & ~size_t(255);
  return UMD_OK;
}

struct Ctx {
  const umd_model_cfg* cfg;
  const umd_step_shape* sh;
  const long long* offs;
  const float* pf;
  const bf16* pb;
  float* grads;
  cudaStream_t st;
  Plan P;
  const float* W(int leaf, long long extra = 0) const { return pf + offs[leaf] + extra; }
  const bf16* WB(int leaf, long long extra = 0) const { return pb + offs[leaf] + extra; }
  float* G(int leaf, long long extra = 0) const { return grads + offs[leaf] + extra; }
  bool has(int leaf) const { return offs[leaf] >= 0; }
};

// The stacks' layer strides travel behind the leaf offsets (include/umd_b200.h: offsets[UMD_P_ENC_LAYER_STRIDE / _DEC_]).
int set_layer_strides(Ctx& c) {
  c.P.enc.lstride = c.offs[UMD_P_ENC_LAYER_STRIDE];
  c.P.dec.lstride = c.offs[UMD_P_DEC_LAYER_STRIDE];
  UMD_REQUIRE(c.P.enc.lstride > 0 && c.P.dec.lstride > 0 && c.P.enc.lstride % 64 == 0 && c.P.dec.lstride % 64 == 0,
              "offsets[UMD_P_ENC_LAYER_STRIDE / UMD_P_DEC_LAYER_STRIDE] must hold the layer strides of the layer-major arena");
  return UMD_OK;
}

umd_gemm_args gemm_base(const void* A, const void* B, int M, int N, int K) {
  umd_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.B = B; g.M = M; g.N = N; g.K = K; g.batch = 1;
  g.split_k = 1;
  g.split_row = M; g.s0 = M > 0 ? M : 1; g.s1 = 1; g.n0 = 1;
  return g;
}
void set_rowmap(umd_gemm_args& g, const RowMap& rm) {
  g.split_row = rm.split_row; g.s0 = rm.s0; g.s1 = rm.s1; g.n0 = rm.n0;
}
int pick_split(int M, int N, int K, int batch) {
  // Split-K factor of a weight-gradient GEMM: the persistent kernel runs ceil(items / SMs) rounds of
  // (k-blocks per item + a fixed fill/drain cost); pick the factor that minimises that product.
  const int bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  const long long tiles = static_cast<long long>(ceil_div(M, 128)) * ceil_div(N, bn) * batch;
  if (g_sm == 0) g_sm = sm_count();
  const int kb = ceil_div(K, 64);
  const int cap = kb / 4 > 0 ? kb / 4 : 1;
  const double fixed = 6.0;  // pipeline fill + accumulator drain, in k-block units
  int best = 1;
  double best_t = 1e30;
  for (int sp = 1; sp <= cap && sp <= 64; ++sp) {
    const long long items = tiles * sp;
    const double rounds = static_cast<double>((items + g_sm - 1) / g_sm);
    const double t = rounds * (static_cast<double>(ceil_div(kb, sp)) + fixed);
    if (t < best_t * 0.999) { best_t = t; best = sp; }
  }
  return best;
}

// Y[M,N] = X[M,K] W[K,N] (+bias) : forward Dense with the Flax [in,out] kernel as an MN-major B operand.
int dense_fwd(const Ctx& c, const bf16* X, int M, int K, const bf16* Wk, int N, const float* bias, int epi, void* out0,
              void* out1 = nullptr, const void* aux = nullptr, const float* gate = nullptr, long long ldgate = 0,
              const RowMap* rm = nullptr) {
  umd_gemm_args g = gemm_base(X, Wk, M, N, K);
  g.a_mn = 0; g.b_mn = 1; g.lda = K; g.ldb = N;
  g.epi = epi; g.out0 = out0; g.ld0 = N; g.out1 = out1; g.ld1 = N; g.bias = bias;
  g.aux = aux; g.ldaux = N; g.gate = gate; g.ldgate = ldgate;
  if (rm) set_rowmap(g, *rm);
  return gemm_bf16(g, c.st);
}
// dX[M,K] = dY[M,N] W[K,N]^T : W in its native layout is a K-major B operand with N_gemm = K, K_gemm = N.
int dense_dgrad(const Ctx& c, const bf16* dY, int M, int N, const bf16* Wk, int K, int epi, void* out0,
                const void* aux = nullptr) {
  umd_gemm_args g = gemm_base(dY, Wk, M, K, N);
  g.a_mn = 0; g.b_mn = 0; g.lda = N; g.ldb = N;
  g.epi = epi; g.out0 = out0; g.ld0 = K; g.aux = aux; g.ldaux = K;
  return gemm_bf16(g, c.st);
}
// dW[K,N] += X[M,K]^T dY[M,N]
int dense_wgrad(const Ctx& c, const bf16* X, int M, int K, const bf16* dY, int N, long long lddy, float* dW) {
  umd_gemm_args g = gemm_base(X, dY, K, N, M);
  g.a_mn = 1; g.b_mn = 1; g.lda = K; g.ldb = lddy;
  g.epi = UMD_EPI_ATOMIC; g.out0 = dW; g.ld0 = N;
  g.split_k = pick_split(K, N, M, 1);
  return gemm_bf16(g, c.st);
}

// ------------------------------------------------------------------------------------------
// conditioning (ae.py:105-124; embeddings.py)
// ------------------------------------------------------------------------------------------
int cond_forward(Ctx& c, const umd_io& io) {
  Plan& P = c.P;
  const int B = P.B, D = P.D;
  UMD_TRY(time_embed(io.t, B, D, P.temb, c.st));
  UMD_TRY(dense_fwd(c, P.temb, B, D, c.WB(UMD_P_TT_W0), 2 * D, c.W(UMD_P_TT_B0), UMD_EPI_F32, P.th1));
  UMD_TRY(silu_cast(P.th1, static_cast<long long>(B) * 2 * D, P.ta1, c.st));
  UMD_TRY(dense_fwd(c, P.ta1, B, 2 * D, c.WB(UMD_P_TT_W1), D, c.W(UMD_P_TT_B1), UMD_EPI_F32, P.tc));
  if (P.has_label) {
    UMD_REQUIRE(io.labels != nullptr, "labels are required when num_classes > 0 (pass the null class id for y=None)");
    UMD_TRY(gather_rows(c.W(UMD_P_LABEL_TABLE), io.labels, B, D, P.lemb, c.st));
    UMD_TRY(dense_fwd(c, P.lemb, B, D, c.WB(UMD_P_LT_W0), 2 * D, c.W(UMD_P_LT_B0), UMD_EPI_F32, P.lh1));
    UMD_TRY(silu_cast(P.lh1, static_cast<long long>(B) * 2 * D, P.la1, c.st));
    UMD_TRY(dense_fwd(c, P.la1, B, 2 * D, c.WB(UMD_P_LT_W1), D, c.W(UMD_P_LT_B1), UMD_EPI_F32, P.yc));
  }
  UMD_TRY(cond_combine(P.tc, P.yc, static_cast<long long>(B) * D, P.adaln, P.s, P.cond, P.cond_bf16, c.st));
  if (P.adaln) {
    // adaLN projections of every block in one batched GEMM per stack (vit.py:71-72), final modulation ae.py:167
    Stack* stacks[2] = {&P.enc, &P.dec};
    for (Stack* s : stacks) {
      umd_gemm_args g = gemm_base(P.cond_bf16, c.WB(s->base + UMD_S_ADA_W), B, 6 * D, D);
      g.b_mn = 1; g.lda = D; g.ldb = 6 * D; g.batch = s->depth; g.a_bs = 0; g.b_bs = s->lstride;
      g.epi = UMD_EPI_F32; g.out0 = s->ada; g.ld0 = 6 * D; g.bs0 = static_cast<long long>(B) * 6 * D;
      g.bias = c.W(s->base + UMD_S_ADA_B); g.bias_bs = s->lstride;
      UMD_TRY(gemm_bf16(g, c.st));
    }
    UMD_TRY(dense_fwd(c, P.cond_bf16, B, D, c.WB(UMD_P_FMOD_W), 2 * D, c.W(UMD_P_FMOD_B), UMD_EPI_F32, P.fmod));
  }
  return UMD_OK;
}

// ------------------------------------------------------------------------------------------
// Encoder1DBlock stack forward (vit.py:60-163)
// ------------------------------------------------------------------------------------------
// The residual update that closes block l, x[l+1] = xmid[l] + gate1[l] * z[l] (vit.py:106-108), is not run as a
// pass of its own: the LayerNorm that consumes x[l+1] performs it on the way in and writes x[l+1] to x_out.
void pending_residual(LnFwdArgs& ln, const Stack& s, int l, void* x_out) {
  const LayerBufs& lb = s.L[l];
  ln.x = lb.xmid; ln.x_bf16 = s.rbf; ln.xout_bf16 = s.rbf;
  ln.res_branch = lb.z;
  ln.res_gate = s.ada ? s.ada + static_cast<long long>(l) * s.nsamples * 6 * s.D + 5 * s.D : nullptr;
  ln.ldgate = 6 * s.D;
  ln.x_out = x_out;
}

int stack_forward(Ctx& c, Stack& s) {
  Plan& P = c.P;
  const int D = P.D, T = s.rows, M4 = P.M4;
  const long long ldada = 6 * D, ada_ls = static_cast<long long>(P.B) * 6 * D, ls = s.lstride;
  const long long qkv_sp = c.offs[s.base + UMD_S_K_W] - c.offs[s.base + UMD_S_Q_W];
  const long long qkvb_sp = c.offs[s.base + UMD_S_K_B] - c.offs[s.base + UMD_S_Q_B];
  UMD_REQUIRE(c.offs[s.base + UMD_S_V_W] - c.offs[s.base + UMD_S_K_W] == qkv_sp &&
                  c.offs[s.base + UMD_S_V_B] - c.offs[s.base + UMD_S_K_B] == qkvb_sp && qkv_sp > 0 && qkvb_sp > 0,
              "query/key/value leaves must be equally spaced in the arena");
  for (int l = 0; l < s.depth; ++l) {
    LayerBufs& lb = s.L[l];
    const float* ada = s.ada ? s.ada + l * ada_ls : nullptr;
    LnFwdArgs ln;
    memset(&ln, 0, sizeof(ln));
    // LayerNorm_0 (+ modulate); for l > 0 it first forms x[l] = xmid[l-1] + gate1[l-1] * z[l-1] (vit.py:106-108),
    // and for adaln=False it plants the conditioning token in row 0 of every sample (vit.py:73-74)
    if (l == 0) {
      ln.x = s.x[0];
    } else {
      pending_residual(ln, s, l - 1, s.x[l]);
    }
    if (!P.adaln) { ln.cond_row = P.cond; ln.x_out = s.x[l]; ln.xout_bf16 = (l > 0) ? s.rbf : 0; }
    ln.gamma = c.W(s.base + UMD_S_LN0_S, l * ls);
    ln.beta = c.W(s.base + UMD_S_LN0_B, l * ls);
    ln.shift = ada; ln.scale = ada ? ada + D : nullptr; ln.ldmod = ldada; ln.rm = s.rm;
    ln.out = lb.y0; ln.mean = lb.mean0; ln.rstd = lb.rstd0; ln.rows_out = T;
    UMD_TRY(ln_mod_fwd(ln, D, true, c.st));
    {  // q, k, v projections as one batch-3 GEMM into the packed [T, 3D] buffer
      umd_gemm_args g = gemm_base(lb.y0, c.WB(s.base + UMD_S_Q_W, l * ls), T, D, D);
      g.b_mn = 1; g.lda = D; g.ldb = D; g.batch = 3; g.a_bs = 0; g.b_bs = qkv_sp;
      g.epi = UMD_EPI_BF16; g.out0 = lb.qkv; g.ld0 = 3 * D; g.bs0 = D;
      g.bias = c.W(s.base + UMD_S_Q_B, l * ls); g.bias_bs = qkvb_sp;
      UMD_TRY(gemm_bf16(g, c.st));
    }
    AttnArgs at;
    at.qkv = lb.qkv; at.out = lb.o; at.lse = lb.lse; at.rm = s.rm; at.nsamples = s.nsamples; at.H = P.H; at.Dh = P.Dh;
    at.scale = 1.0f / sqrtf(static_cast<float>(P.Dh));
    UMD_TRY(attention_fwd(at, c.st));
    UMD_TRY(dense_fwd(c, lb.o, T, D, c.WB(s.base + UMD_S_O_W, l * ls), D, c.W(s.base + UMD_S_O_B, l * ls), UMD_EPI_BF16, lb.a));
    // LayerNorm_1 (+ modulate) on xmid = x[l] + gate0 * a (vit.py:89-98)
    memset(&ln, 0, sizeof(ln));
    ln.x = s.x[l]; ln.x_bf16 = (l > 0) ? s.rbf : 0;
    ln.res_branch = lb.a; ln.res_gate = ada ? ada + 2 * D : nullptr; ln.ldgate = ldada; ln.x_out = lb.xmid; ln.xout_bf16 = s.rbf;
    ln.gamma = c.W(s.base + UMD_S_LN1_S, l * ls);
    ln.beta = c.W(s.base + UMD_S_LN1_B, l * ls);
    ln.shift = ada ? ada + 3 * D : nullptr; ln.scale = ada ? ada + 4 * D : nullptr; ln.ldmod = ldada; ln.rm = s.rm;
    ln.out = lb.y1; ln.mean = lb.mean1; ln.rstd = lb.rstd1; ln.rows_out = T;
    UMD_TRY(ln_mod_fwd(ln, D, true, c.st));
    UMD_TRY(dense_fwd(c, lb.y1, T, D, c.WB(s.base + UMD_S_FC1_W, l * ls), M4, c.W(s.base + UMD_S_FC1_B, l * ls), UMD_EPI_GELU,
                      lb.u, lb.g));
    UMD_TRY(dense_fwd(c, lb.g, T, M4, c.WB(s.base + UMD_S_FC2_W, l * ls), D, c.W(s.base + UMD_S_FC2_B, l * ls), UMD_EPI_BF16,
                      lb.z));
    // x[l+1] = xmid + gate1 * z is formed by whichever LayerNorm reads it next (pending_residual)
  }
  return UMD_OK;
}

// Gate backward of the MLP branch of block l (App. E step 1), run as the tail of the LayerNorm backward that
// finalises d x[l+1]:  dzb = gate1 * dx, dgate1 = sum_t z dx, d fc2-bias += sum_t dzb.
void gate_stage_mlp(const Ctx& c, const Stack& s, int l, LnBwdArgs& lnb) {
  const Plan& P = c.P;
  const int D = P.D;
  const long long ldada = 6 * D, ada_ls = static_cast<long long>(P.B) * 6 * D;
  const float* ada = s.ada ? s.ada + l * ada_ls : nullptr;
  float* dada = s.dada ? s.dada + l * ada_ls : nullptr;
  lnb.g_dz = P.dzb; lnb.g_z = s.L[l].z;
  lnb.g_gate = ada ? ada + 5 * D : nullptr; lnb.g_ldgate = ldada;
  lnb.g_dgate = dada ? dada + 5 * D : nullptr; lnb.g_lddgate = ldada;
  lnb.g_dbias = c.G(s.base + UMD_S_FC2_B, l * s.lstride);
}

// Backward of the stack (App. E steps 1-10).  On entry dx holds d x[depth] and P.dzb the gated gradient of the
// last block's MLP branch (gate_stage_mlp fused into the caller's LayerNorm backward); on exit dx holds d x[0].
// Every parameter gradient of layer l (the adaLN projection included) is final once layer l's kernels have been
// enqueued; then cb(cb_user, ev0 + depth-1-l) tells the host that the layer's arena block may be all-reduced.
// dxb != null: the running gradient stream is bf16 in dxb; the last LayerNorm backward of the stack (layer 0) reads it and
// writes the stack's input gradient to dx in fp32 (its consumers, the embedding / decoder-input backward, read fp32).
int stack_backward(Ctx& c, Stack& s, float* dx, bf16* dxb, umd_bucket_cb cb, void* cb_user, int ev0) {
  auto set_dx = [&](LnBwdArgs& lnb, bool last) {
    if (dxb) {
      lnb.dx_in = dxb; lnb.dxin_bf16 = 1;
      if (last) { lnb.dx = dx; lnb.dxout_bf16 = 0; } else { lnb.dx = dxb; lnb.dxout_bf16 = 1; }
    } else {
      lnb.dx = dx; lnb.dx_in = dx; lnb.dxin_bf16 = 0; lnb.dxout_bf16 = 0;
    }
    lnb.accumulate = 1;
  };
  Plan& P = c.P;
  const int D = P.D, T = s.rows, M4 = P.M4, B = P.B;
  const long long ldada = 6 * D, ada_ls = static_cast<long long>(B) * 6 * D, ls = s.lstride;
  const long long qkv_sp = c.offs[s.base + UMD_S_K_W] - c.offs[s.base + UMD_S_Q_W];
  const long long qkvb_sp = c.offs[s.base + UMD_S_K_B] - c.offs[s.base + UMD_S_Q_B];
  for (int l = s.depth - 1; l >= 0; --l) {
    LayerBufs& lb = s.L[l];
    const float* ada = s.ada ? s.ada + l * ada_ls : nullptr;
    float* dada = s.dada ? s.dada + l * ada_ls : nullptr;
    const long long lo = l * ls;   // arena offset of this layer's block relative to layer 0
    // ---- MLP branch (P.dzb = gate1 * dx)
    UMD_TRY(dense_dgrad(c, P.dzb, T, D, c.WB(s.base + UMD_S_FC2_W, lo), M4, UMD_EPI_DGELU, P.dgb, lb.u));
    UMD_TRY(dense_wgrad(c, lb.g, T, M4, P.dzb, D, D, c.G(s.base + UMD_S_FC2_W, lo)));
    cudaEvent_t side_a = nullptr, side_b = nullptr;
    UMD_TRY(on_side(c.st, &side_a, [&](cudaStream_t st) {   // d fc1-bias = column sums of dU, behind the next GEMMs
      return colsum_bf16(P.dgb, M4, T, M4, c.G(s.base + UMD_S_FC1_B, lo), st);
    }));
    UMD_TRY(dense_dgrad(c, P.dgb, T, M4, c.WB(s.base + UMD_S_FC1_W, lo), D, UMD_EPI_BF16, P.dyb));
    UMD_TRY(dense_wgrad(c, lb.y1, T, D, P.dgb, M4, M4, c.G(s.base + UMD_S_FC1_W, lo)));
    LnBwdArgs lnb;
    memset(&lnb, 0, sizeof(lnb));
    lnb.dy = P.dyb; lnb.x = lb.xmid; lnb.x_bf16 = s.rbf; lnb.mean = lb.mean1; lnb.rstd = lb.rstd1;
    lnb.gamma = c.W(s.base + UMD_S_LN1_S, lo); lnb.beta = c.W(s.base + UMD_S_LN1_B, lo);
    lnb.scale = ada ? ada + 4 * D : nullptr; lnb.ldmod = ldada; lnb.rm = s.rm; set_dx(lnb, false);
    lnb.dshift = dada ? dada + 3 * D : nullptr; lnb.dscale = dada ? dada + 4 * D : nullptr; lnb.ldd = ldada;
    lnb.dgamma = c.G(s.base + UMD_S_LN1_S, lo); lnb.dbeta = c.G(s.base + UMD_S_LN1_B, lo);
    // ... followed in the same pass by the gate backward of the attention branch (App. E step 6)
    lnb.g_dz = P.dzb; lnb.g_z = lb.a; lnb.g_gate = ada ? ada + 2 * D : nullptr; lnb.g_ldgate = ldada;
    lnb.g_dgate = dada ? dada + 2 * D : nullptr; lnb.g_lddgate = ldada; lnb.g_dbias = c.G(s.base + UMD_S_O_B, lo);
    UMD_TRY(ln_mod_bwd(lnb, D, s.nsamples, true, c.st));
    // ---- attention branch (P.dzb = gate0 * dx)
    {  // dO = dA Wo^T, with delta = rowsum(dO o O) per head from the same accumulators (App. E step 7): the attention
       // backward then needs neither O nor a prologue of its own
      umd_gemm_args g = gemm_base(P.dzb, c.WB(s.base + UMD_S_O_W, lo), T, D, D);
      g.a_mn = 0; g.b_mn = 0; g.lda = D; g.ldb = D;
      g.epi = UMD_EPI_BF16_DELTA; g.out0 = P.dyb; g.ld0 = D; g.aux = lb.o; g.ldaux = D; g.out1 = P.delta; g.ld1 = P.H;
      UMD_TRY(gemm_bf16(g, c.st));
    }
    UMD_TRY(dense_wgrad(c, lb.o, T, D, P.dzb, D, D, c.G(s.base + UMD_S_O_W, lo)));
    AttnBwdArgs ab;
    ab.qkv = lb.qkv; ab.out = nullptr; ab.delta = P.delta; ab.dout = P.dyb; ab.lse = lb.lse; ab.dqkv = P.dqkv; ab.rm = s.rm;
    ab.nsamples = s.nsamples; ab.H = P.H; ab.Dh = P.Dh; ab.scale = 1.0f / sqrtf(static_cast<float>(P.Dh));
    UMD_TRY(attention_bwd(ab, c.st));
    {  // dWq, dWk, dWv as one batch-3 wgrad GEMM
      umd_gemm_args g = gemm_base(lb.y0, P.dqkv, D, D, T);
      g.a_mn = 1; g.b_mn = 1; g.lda = D; g.ldb = 3 * D; g.batch = 3; g.a_bs = 0; g.b_bs = D;
      g.epi = UMD_EPI_ATOMIC; g.out0 = c.G(s.base + UMD_S_Q_W, lo); g.ld0 = D; g.bs0 = qkv_sp;
      g.split_k = pick_split(D, D, T, 3);
      UMD_TRY(gemm_bf16(g, c.st));
    }
    UMD_TRY(on_side(c.st, &side_b, [&](cudaStream_t st) {   // dbq | dbk | dbv
      return colsum_bf16(P.dqkv, 3 * D, T, 3 * D, c.G(s.base + UMD_S_Q_B, lo), st, D, qkvb_sp);
    }));
    {  // dY0 = [dQ|dK|dV] [Wq|Wk|Wv]^T, contraction chunked over the three kernels
      umd_gemm_args g = gemm_base(P.dqkv, c.WB(s.base + UMD_S_Q_W, lo), T, D, 3 * D);
      g.a_mn = 0; g.b_mn = 0; g.lda = 3 * D; g.ldb = D; g.b_bs = qkv_sp; g.b_kchunk = D;
      g.epi = UMD_EPI_BF16; g.out0 = P.dyb; g.ld0 = D;
      UMD_TRY(gemm_bf16(g, c.st));
    }
    memset(&lnb, 0, sizeof(lnb));
    lnb.dy = P.dyb; lnb.x = s.x[l]; lnb.x_bf16 = (l > 0) ? s.rbf : 0; lnb.mean = lb.mean0; lnb.rstd = lb.rstd0;
    lnb.gamma = c.W(s.base + UMD_S_LN0_S, lo); lnb.beta = c.W(s.base + UMD_S_LN0_B, lo);
    lnb.scale = ada ? ada + D : nullptr; lnb.ldmod = ldada; lnb.rm = s.rm; set_dx(lnb, l == 0);
    lnb.dshift = dada ? dada : nullptr; lnb.dscale = dada ? dada + D : nullptr; lnb.ldd = ldada;
    lnb.dgamma = c.G(s.base + UMD_S_LN0_S, lo); lnb.dbeta = c.G(s.base + UMD_S_LN0_B, lo);
    if (!P.adaln) lnb.dcond = P.dcond;            // token-0 row: gradient of the conditioning token (vit.py:73-74)
    if (l > 0) gate_stage_mlp(c, s, l - 1, lnb);  // dx is now d x[l]: start block l-1's MLP-branch backward
    UMD_TRY(ln_mod_bwd(lnb, D, s.nsamples, true, c.st));
    if (P.adaln) {
      // adaLN projection of this layer (App. E step 10): dada[l] = [dshift0|dscale0|dgate0|dshift1|dscale1|dgate1] per
      // sample is complete (the gate1 column was written by layer l+1's... no: by THIS layer's first LayerNorm backward of
      // the caller / of layer l+1, all earlier in the stream)
      bf16* dab = s.dada_bf16 + l * ada_ls;
      UMD_TRY(cast_bf16(dada, ada_ls, dab, c.st));
      UMD_TRY(colsum_f32(dada, ldada, B, 6 * D, c.G(s.base + UMD_S_ADA_B, lo), c.st));
      UMD_TRY(dense_wgrad(c, P.cond_bf16, B, D, dab, 6 * D, ldada, c.G(s.base + UMD_S_ADA_W, lo)));
    }
    // the side passes read P.dgb / P.dqkv (rewritten by the next layer) and write this layer's bias gradients
    UMD_TRY(join_side(c.st, side_a));
    UMD_TRY(join_side(c.st, side_b));
    if (cb) cb(cb_user, ev0 + (s.depth - 1 - l));
  }
  if (P.adaln) {
    // dcond += dada[l] W_ada[l]^T for all blocks of the stack in one batched GEMM (it feeds no parameter gradient
    // directly, so it can wait for the end of the stack)
    umd_gemm_args g = gemm_base(s.dada_bf16, c.WB(s.base + UMD_S_ADA_W), B, D, 6 * D);
    g.a_mn = 0; g.b_mn = 0; g.lda = ldada; g.ldb = 6 * D; g.batch = s.depth; g.a_bs = ada_ls; g.b_bs = ls;
    g.epi = UMD_EPI_ATOMIC; g.out0 = P.dcond; g.ld0 = D; g.bs0 = 0;
    UMD_TRY(gemm_bf16(g, c.st));
  }
  return UMD_OK;
}

EmbedArgs embed_args(const Ctx& c, const umd_io& io) {
  const Plan& P = c.P;
  EmbedArgs e;
  e.image = io.image; e.ids_keep = io.ids_shuffle; e.W = c.W(UMD_P_EMBED_W); e.bias = c.W(UMD_P_EMBED_B);
  e.pos = c.W(UMD_P_POS); e.cls = c.W(UMD_P_CLS); e.x = static_cast<float*>(P.enc.x[0]); e.rm = P.enc.rm; e.n1 = c.sh->n1;
  e.keep0 = c.sh->keep0; e.keep1 = c.sh->keep1; e.masked0 = c.sh->masked0; e.masked1 = c.sh->masked1;
  e.img = c.cfg->img_size; e.patch = P.p; e.C = P.C; e.D = P.D; e.L = P.L; e.num_cls = P.ncls; e.tok0 = P.tok0;
  return e;
}
DecInArgs decin_args(const Ctx& c, const umd_io& io) {
  const Plan& P = c.P;
  DecInArgs d;
  d.enc = P.enc.xf; d.ids_restore = io.ids_restore; d.ids_keep = io.ids_shuffle; d.mask_token = c.W(UMD_P_MASK_TOKEN);
  d.dec_pos = c.W(UMD_P_DEC_POS); d.xd = static_cast<float*>(P.dec.x[0]); d.rep = io.pre_logits; d.rm_enc = P.enc.rm;
  d.keep0 = c.sh->keep0; d.keep1 = c.sh->keep1; d.masked0 = c.sh->masked0; d.masked1 = c.sh->masked1;
  d.D = P.D; d.L = P.L; d.num_cls = P.ncls; d.tok0 = P.tok0; d.S_d = P.Sd;
  return d;
}
LossArgs loss_args(const Ctx& c, const umd_io& io, bool with_grad) {
  const Plan& P = c.P;
  const umd_step_shape& sh = *c.sh;
  LossArgs a;
  a.predp = P.predp; a.x0 = io.x0; a.noise = io.noise; a.ids_restore = io.ids_restore;
  a.dpredp = with_grad ? P.dpredp : nullptr; a.partials = P.loss_partials;
  a.n0 = sh.n0; a.n1 = sh.n1; a.keep0 = sh.keep0; a.keep1 = sh.keep1; a.masked0 = sh.masked0; a.masked1 = sh.masked1;
  a.img = c.cfg->img_size; a.patch = P.p; a.C = P.C; a.L = P.L; a.grad_scale = 1.f;
  // train_ae.py:333-335,352-360: loss = dit*(1 - n1/B) + mae*(n1/B); each mean()/mean(mask) ratio reduces to a
  // sum divided by C * (number of selected pixels) because every sample masks the same number of patches.
  const double Bt = sh.n0 + sh.n1, pp = static_cast<double>(P.p) * P.p, C = P.C;
  const double wn = 1.0 - sh.n1 / Bt, wm = sh.n1 / Bt;
  a.w_x0_0 = a.w_eps_0 = a.w_x0_1 = 0.f;
  if (sh.n0 > 0) {
    if (sh.masked0) {
      a.w_x0_0 = static_cast<float>(0.5 * wn / (C * sh.n0 * (P.L - sh.keep0) * pp));
      a.w_eps_0 = static_cast<float>(0.5 * wn / (C * sh.n0 * sh.keep0 * pp));
    } else {
      a.w_x0_0 = a.w_eps_0 = static_cast<float>(0.5 * wn / (C * sh.n0 * P.L * pp));
    }
  }
  if (sh.n1 > 0) a.w_x0_1 = static_cast<float>(wm / (C * sh.n1 * (P.L - sh.keep1) * pp));
  return a;
}

int check_common(const umd_model_cfg* cfg, const umd_step_shape* shape, const long long* offsets, const void* params,
                 const void* params_bf16, const umd_io* io, void* ws) {
  UMD_REQUIRE(cfg && shape && offsets && params && params_bf16 && io && ws, "null argument");
  UMD_REQUIRE(io->image && io->t, "image and t are required");
  UMD_REQUIRE(!((shape->masked0 && shape->n0 > 0) || (shape->masked1 && shape->n1 > 0)) ||
                  (io->ids_shuffle && io->ids_restore),
              "ids_shuffle / ids_restore are required for masked segments");
  UMD_REQUIRE(shape->n1 == 0 || shape->masked1, "the clean (MAE) segment must be masked (train_ae.py:335 divides by mean(mask))");
  return UMD_OK;
}

}  // namespace

int engine_forward(const umd_model_cfg* cfg, const umd_step_shape* shape, const long long* offsets, const float* params,
                   const void* params_bf16, const umd_io* io, void* ws, size_t ws_bytes, int train, cudaStream_t st) {
  UMD_TRY(check_common(cfg, shape, offsets, params, params_bf16, io, ws));
  Ctx c;
  c.cfg = cfg; c.sh = shape; c.offs = offsets; c.pf = params; c.pb = static_cast<const bf16*>(params_bf16);
  c.grads = nullptr; c.st = st;
  UMD_TRY(make_plan(c.P, *cfg, *shape, ws, train != 0));
  UMD_REQUIRE(c.P.bytes <= ws_bytes, "workspace too small: need %zu bytes, have %zu", c.P.bytes, ws_bytes);
  UMD_TRY(set_layer_strides(c));
  Plan& P = c.P;
  const int D = P.D, B = P.B;
  UMD_TRY(cond_forward(c, *io));
  UMD_TRY(embed_fwd(embed_args(c, *io), B, st));
  UMD_TRY(stack_forward(c, P.enc));
  {  // encoder_norm (vit.py:163), fp32 out
    LnFwdArgs ln;
    memset(&ln, 0, sizeof(ln));
    pending_residual(ln, P.enc, P.enc.depth - 1, P.enc.x[P.enc.depth]);
    ln.gamma = c.W(UMD_P_ENC_BASE + UMD_S_NORM_S); ln.beta = c.W(UMD_P_ENC_BASE + UMD_S_NORM_B);
    ln.rm = P.enc.rm; ln.out = P.enc.xf; ln.mean = P.enc.meanf; ln.rstd = P.enc.rstdf; ln.rows_out = P.Te;
    UMD_TRY(ln_mod_fwd(ln, D, false, st));
  }
  UMD_TRY(decoder_input_fwd(decin_args(c, *io), B, st));
  UMD_TRY(stack_forward(c, P.dec));
  {  // decoder encoder_norm + drop the rep row + final modulation (ae.py:163-170) -> bf16 GEMM operand
    LnFwdArgs ln;
    memset(&ln, 0, sizeof(ln));
    pending_residual(ln, P.dec, P.dec.depth - 1, P.dec.x[P.dec.depth]);
    ln.gamma = c.W(UMD_P_DEC_BASE + UMD_S_NORM_S); ln.beta = c.W(UMD_P_DEC_BASE + UMD_S_NORM_B);
    ln.shift = P.fmod; ln.scale = P.fmod ? P.fmod + D : nullptr; ln.ldmod = 2 * D; ln.rm = P.dec.rm;
    ln.out = P.xm; ln.mean = P.meanF; ln.rstd = P.rstdF; ln.rows_out = B * P.L; ln.gather_L = P.L; ln.gather_off = P.tok0 + 1;
    UMD_TRY(ln_mod_fwd(ln, D, true, st));
  }
  UMD_TRY(pack_final_conv(c.W(UMD_P_FCONV_W), c.W(UMD_P_FCONV_B), P.p, D, 2 * P.C, cfg->flip_final_conv, P.Wf_mat, P.biasm, st));
  UMD_TRY(dense_fwd(c, P.xm, B * P.L, D, P.Wf_mat, P.NC, P.biasm, UMD_EPI_F32, P.predp));
  if (io->pred) UMD_TRY(unpatchify(P.predp, B, cfg->img_size, P.p, 2 * P.C, io->pred, st));
  if (io->x0 && io->loss) {
    UMD_REQUIRE(shape->n0 == 0 || io->noise, "noise targets are required for the noise segment");
    UMD_TRY(loss_fwd_bwd(loss_args(c, *io, train != 0), io->loss, st));
  }
  return UMD_OK;
}

int engine_backward(const umd_model_cfg* cfg, const umd_step_shape* shape, const long long* offsets, const float* params,
                    const void* params_bf16, float* grads, const umd_io* io, void* ws, size_t ws_bytes, umd_bucket_cb cb,
                    void* cb_user, cudaStream_t st) {
  UMD_TRY(check_common(cfg, shape, offsets, params, params_bf16, io, ws));
  UMD_REQUIRE(grads != nullptr, "grads is null");
  Ctx c;
  c.cfg = cfg; c.sh = shape; c.offs = offsets; c.pf = params; c.pb = static_cast<const bf16*>(params_bf16);
  c.grads = grads; c.st = st;
  UMD_TRY(make_plan(c.P, *cfg, *shape, ws, true));
  UMD_REQUIRE(c.P.bytes <= ws_bytes, "workspace too small: need %zu bytes, have %zu", c.P.bytes, ws_bytes);
  UMD_TRY(set_layer_strides(c));
  Plan& P = c.P;
  const int D = P.D, B = P.B, BL = B * P.L;
  UMD_CHECK_CUDA(cudaMemsetAsync(P.dcond, 0, static_cast<size_t>(B) * D * sizeof(float), st));
  UMD_CHECK_CUDA(cudaMemsetAsync(P.dWf_mat, 0, static_cast<size_t>(D) * P.NC * sizeof(float), st));
  UMD_CHECK_CUDA(cudaMemsetAsync(P.dbiasm, 0, static_cast<size_t>(P.NC) * sizeof(float), st));
  if (P.adaln) {  // per-sample modulation gradients are accumulated atomically by the LayerNorm backward kernels
    UMD_CHECK_CUDA(cudaMemsetAsync(P.enc.dada, 0, static_cast<size_t>(B) * P.enc.depth * 6 * D * sizeof(float), st));
    UMD_CHECK_CUDA(cudaMemsetAsync(P.dec.dada, 0, static_cast<size_t>(B) * P.dec.depth * 6 * D * sizeof(float), st));
    UMD_CHECK_CUDA(cudaMemsetAsync(P.dfmod, 0, static_cast<size_t>(B) * 2 * D * sizeof(float), st));
  }
  // ---- un-patchify (final_conv) backward
  UMD_TRY(dense_wgrad(c, P.xm, BL, D, P.dpredp, P.NC, P.NC, P.dWf_mat));
  UMD_TRY(colsum_bf16(P.dpredp, P.NC, BL, P.NC, P.dbiasm, st));
  UMD_TRY(unpack_final_conv_grad(P.dWf_mat, P.dbiasm, P.p, D, 2 * P.C, cfg->flip_final_conv, c.G(UMD_P_FCONV_W),
                                 c.G(UMD_P_FCONV_B), st));
  UMD_TRY(dense_dgrad(c, P.dpredp, BL, P.NC, P.Wf_mat, D, UMD_EPI_BF16, P.dyb));
  {  // decoder encoder_norm + final modulation backward
    LnBwdArgs lnb;
    memset(&lnb, 0, sizeof(lnb));
    lnb.dy = P.dyb; lnb.x = P.dec.x[P.dec.depth]; lnb.x_bf16 = P.dec.rbf; lnb.mean = P.meanF; lnb.rstd = P.rstdF;
    lnb.gamma = c.W(UMD_P_DEC_BASE + UMD_S_NORM_S); lnb.beta = c.W(UMD_P_DEC_BASE + UMD_S_NORM_B);
    lnb.scale = P.fmod ? P.fmod + D : nullptr; lnb.ldmod = 2 * D; lnb.rm = P.dec.rm; lnb.accumulate = 0;
    if (P.dxb_dec) { lnb.dx = P.dxb_dec; lnb.dxout_bf16 = 1; } else { lnb.dx = P.dx_dec; }
    lnb.dx_in = lnb.dx;
    lnb.dshift = P.dfmod; lnb.dscale = P.dfmod ? P.dfmod + D : nullptr; lnb.ldd = 2 * D;
    lnb.dgamma = c.G(UMD_P_DEC_BASE + UMD_S_NORM_S); lnb.dbeta = c.G(UMD_P_DEC_BASE + UMD_S_NORM_B);
    lnb.gather_L = P.L; lnb.gather_off = P.tok0 + 1;
    gate_stage_mlp(c, P.dec, P.dec.depth - 1, lnb);
    UMD_TRY(ln_mod_bwd(lnb, D, B, true, st));
  }
  if (P.adaln) {  // final_modulation Dense backward
    UMD_TRY(cast_bf16(P.dfmod, static_cast<long long>(B) * 2 * D, P.dfmod_bf16, st));
    UMD_TRY(colsum_f32(P.dfmod, 2 * D, B, 2 * D, c.G(UMD_P_FMOD_B), st));
    UMD_TRY(dense_wgrad(c, P.cond_bf16, B, D, P.dfmod_bf16, 2 * D, 2 * D, c.G(UMD_P_FMOD_W)));
    umd_gemm_args g = gemm_base(P.dfmod_bf16, c.WB(UMD_P_FMOD_W), B, D, 2 * D);
    g.lda = 2 * D; g.ldb = 2 * D; g.epi = UMD_EPI_ATOMIC; g.out0 = P.dcond; g.ld0 = D;
    UMD_TRY(gemm_bf16(g, st));
  }
  UMD_TRY(stack_backward(c, P.dec, P.dx_dec, P.dxb_dec, nullptr, nullptr, 0));
  UMD_TRY(decoder_input_bwd(decin_args(c, *io), B, P.Te, P.dx_dec, P.dxf_enc, c.G(UMD_P_DEC_POS), c.G(UMD_P_MASK_TOKEN), st));
  if (cb) cb(cb_user, 0);
  {  // encoder_norm backward
    LnBwdArgs lnb;
    memset(&lnb, 0, sizeof(lnb));
    lnb.dy = P.dxf_enc; lnb.x = P.enc.x[P.enc.depth]; lnb.x_bf16 = P.enc.rbf; lnb.mean = P.enc.meanf; lnb.rstd = P.enc.rstdf;
    lnb.gamma = c.W(UMD_P_ENC_BASE + UMD_S_NORM_S); lnb.beta = c.W(UMD_P_ENC_BASE + UMD_S_NORM_B);
    lnb.rm = P.enc.rm; lnb.accumulate = 0;
    if (P.dxb_enc) { lnb.dx = P.dxb_enc; lnb.dxout_bf16 = 1; } else { lnb.dx = P.dx_enc; }
    lnb.dx_in = lnb.dx;
    lnb.dgamma = c.G(UMD_P_ENC_BASE + UMD_S_NORM_S); lnb.dbeta = c.G(UMD_P_ENC_BASE + UMD_S_NORM_B);
    gate_stage_mlp(c, P.enc, P.enc.depth - 1, lnb);
    UMD_TRY(ln_mod_bwd(lnb, D, B, false, st));
  }
  UMD_TRY(stack_backward(c, P.enc, P.dx_enc, P.dxb_enc, cb, cb_user, 1));
  UMD_TRY(embed_bwd(embed_args(c, *io), B, P.dx_enc, c.G(UMD_P_EMBED_W), c.G(UMD_P_EMBED_B), c.G(UMD_P_POS),
                    c.G(UMD_P_CLS), st));
  // ---- conditioning path backward (ae.py:121-124, embeddings.py:50-59)
  UMD_TRY(cond_combine_bwd(P.dcond, P.s, static_cast<long long>(B) * D, P.adaln, P.ds, st));
  struct Trunk { int w0, b0, w1, b1; const bf16* in; const float* h1; const bf16* a1; bf16* da1; bf16* dh1; };
  Trunk trunks[2] = {{UMD_P_TT_W0, UMD_P_TT_B0, UMD_P_TT_W1, UMD_P_TT_B1, P.temb, P.th1, P.ta1, P.dta1, P.dth1},
                     {UMD_P_LT_W0, UMD_P_LT_B0, UMD_P_LT_W1, UMD_P_LT_B1, P.lemb, P.lh1, P.la1, P.dla1, P.dlh1}};
  for (int k = 0; k < (P.has_label ? 2 : 1); ++k) {
    const Trunk& t = trunks[k];
    UMD_TRY(colsum_bf16(P.ds, D, B, D, c.G(t.b1), st));
    UMD_TRY(dense_wgrad(c, t.a1, B, 2 * D, P.ds, D, D, c.G(t.w1)));
    UMD_TRY(dense_dgrad(c, P.ds, B, D, c.WB(t.w1), 2 * D, UMD_EPI_BF16, t.da1));
    UMD_TRY(silu_bwd(t.da1, t.h1, static_cast<long long>(B) * 2 * D, t.dh1, st));
    UMD_TRY(colsum_bf16(t.dh1, 2 * D, B, 2 * D, c.G(t.b0), st));
    UMD_TRY(dense_wgrad(c, t.in, B, D, t.dh1, 2 * D, 2 * D, c.G(t.w0)));
    if (k == 1) {  // label embedding table (embeddings.py:47)
      umd_gemm_args g = gemm_base(t.dh1, c.WB(t.w0), B, D, 2 * D);
      g.lda = 2 * D; g.ldb = 2 * D; g.epi = UMD_EPI_F32; g.out0 = P.dlemb; g.ld0 = D;
      UMD_TRY(gemm_bf16(g, st));
      UMD_TRY(scatter_add_rows(P.dlemb, io->labels, B, D, c.G(UMD_P_LABEL_TABLE), st));
    }
  }
  if (cb) cb(cb_user, 1 + P.enc.depth);
  return UMD_OK;
}

}  // namespace umd

using namespace umd;

extern "C" size_t umd_workspace_bytes(const umd_model_cfg* cfg, const umd_step_shape* shape, int train) {
  if (!cfg || !shape) return 0;
  Plan P;
  if (make_plan(P, *cfg, *shape, nullptr, train != 0) != UMD_OK) return 0;
  return P.bytes;
}
extern "C" int umd_forward(const umd_model_cfg* cfg, const umd_step_shape* shape, const long long* offsets,
                           const float* params, const void* params_bf16, const umd_io* io, void* workspace,
                           size_t workspace_bytes, int train, umd_stream_t stream) {
  return engine_forward(cfg, shape, offsets, params, params_bf16, io, workspace, workspace_bytes, train,
                        static_cast<cudaStream_t>(stream));
}
extern "C" int umd_backward(const umd_model_cfg* cfg, const umd_step_shape* shape, const long long* offsets,
                            const float* params, const void* params_bf16, float* grads, const umd_io* io, void* workspace,
                            size_t workspace_bytes, umd_bucket_cb cb, void* cb_user, umd_stream_t stream) {
  return engine_backward(cfg, shape, offsets, params, params_bf16, grads, io, workspace, workspace_bytes, cb, cb_user,
                         static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------
// update_fn in one call (train_ae.py:287-382)
// ------------------------------------------------------------------------------------------
namespace umd {
namespace {
struct StepScratch {
  float* model_in;
  int* t_model;
  int* labels_model;
  int* ids_shuffle;
  int* ids_restore;
  size_t bytes;   // appended to the engine workspace
};
StepScratch carve_step(const umd_model_cfg& c, const umd_step_shape& sh, void* base) {
  Bump b(base);
  const long long B = sh.n0 + sh.n1, L = static_cast<long long>(c.img_size / c.patch) * (c.img_size / c.patch);
  StepScratch s;
  s.model_in = b.take<float>(B * c.img_size * c.img_size * c.channels);
  s.t_model = b.take<int>(B);
  s.labels_model = b.take<int>(B);
  s.ids_shuffle = b.take<int>(B * L);
  s.ids_restore = b.take<int>(B * L);
  s.bytes = (b.off + 255) & ~size_t(255);
  return s;
}
struct ReduceCtx {
  const umd_train_step_args* a;
  Comm* comm;
  cudaStream_t st;
  int rc;
};
void reduce_cb(void* user, int event) {
  ReduceCtx* r = static_cast<ReduceCtx*>(user);
  if (r->rc != UMD_OK) return;
  const umd_train_step_args& a = *r->a;
  for (int k = 0; k < a.num_buckets; ++k) {
    if (a.bucket_events[k] != event) continue;
    long long lo = a.bucket_bounds[2 * k], hi = a.bucket_bounds[2 * k + 1];
    if (hi == a.opt.n) hi += 64;   // the trailing scalar slots (loss) ride on the bucket that ends the arena
    const int rc = comm_allreduce_mean_after(r->comm, a.grads + lo, hi - lo, r->st, nullptr);
    if (rc != UMD_OK) { r->rc = rc; return; }
  }
}
}  // namespace
}  // namespace umd

extern "C" size_t umd_train_workspace_bytes(const umd_model_cfg* cfg, const umd_step_shape* shape) {
  const size_t w = umd_workspace_bytes(cfg, shape, 1);
  if (w == 0) return 0;
  return w + carve_step(*cfg, *shape, nullptr).bytes;
}

extern "C" int umd_train_step(const umd_train_step_args* a, umd_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UMD_REQUIRE(a && a->cfg && a->offsets && a->workspace, "umd_train_step: null argument");
  const umd_model_cfg& c = *a->cfg;
  const umd_step_shape& sh = a->shape;
  const int B = sh.n0 + sh.n1;
  UMD_REQUIRE(B > 0 && a->image && a->opt.params && a->grads && a->opt.params_bf16 && a->opt.n > 0,
              "umd_train_step: image, opt.params, opt.params_bf16 and grads are required");
  UMD_REQUIRE(sh.n0 == 0 || (a->t && a->noise && a->sqrt_alphas_cumprod && a->sqrt_one_minus_alphas_cumprod),
              "umd_train_step: t, noise and the schedule tables are required for the noise branch");
  UMD_REQUIRE(!(sh.masked0 && sh.n0 > 0) || a->mask_noise0, "umd_train_step: mask_noise0 is required (segment 0 is masked)");
  UMD_REQUIRE(!(sh.masked1 && sh.n1 > 0) || a->mask_noise1, "umd_train_step: mask_noise1 is required (segment 1 is masked)");
  UMD_REQUIRE(!a->comm || (a->num_buckets > 0 && a->bucket_bounds && a->bucket_events), "umd_train_step: a communicator needs the bucket table");
  const bool optimise = !(a->flags & UMD_STEP_NO_OPTIMIZER);
  UMD_REQUIRE(!optimise || (a->opt.mu && a->opt.nu && a->opt.wd_flags && a->opt.scratch && a->opt.measurements),
              "umd_train_step: the optimiser needs mu, nu, wd_flags, scratch and measurements");
  const size_t ws_engine = umd_workspace_bytes(&c, &sh, 1);
  UMD_REQUIRE(ws_engine > 0, "umd_train_step: %s", umd_last_error());
  StepScratch ss = carve_step(c, sh, static_cast<uint8_t*>(a->workspace) + ws_engine);
  UMD_REQUIRE(ws_engine + ss.bytes <= a->workspace_bytes, "umd_train_step: workspace too small: need %zu bytes, have %zu",
              ws_engine + ss.bytes, a->workspace_bytes);
  const int L = (c.img_size / c.patch) * (c.img_size / c.patch);
  const long long per = static_cast<long long>(c.img_size) * c.img_size * c.channels;
  float* grads = a->grads;
  float* loss = grads + a->opt.n;
  UMD_CHECK_CUDA(cudaMemsetAsync(grads, 0, static_cast<size_t>(a->opt.n + 64) * sizeof(float), st));
  // ---- model inputs: x_t for the noise branch (q_sample, :318-321), x_0 for the clean branch
  if (sh.n0 > 0)
    UMD_TRY(qsample(a->image, a->noise, a->t, a->sqrt_alphas_cumprod, a->sqrt_one_minus_alphas_cumprod, sh.n0,
                    static_cast<int>(per), ss.model_in, st));
  if (sh.n1 > 0)
    UMD_CHECK_CUDA(cudaMemcpyAsync(ss.model_in + sh.n0 * per, a->image + sh.n0 * per, static_cast<size_t>(sh.n1) * per * sizeof(float),
                                   cudaMemcpyDeviceToDevice, st));
  UMD_TRY(step_prep(a->t, a->label, a->label_drop, sh.n0, B, c.num_classes, a->use_labels, ss.t_model,
                    c.num_classes > 0 ? ss.labels_model : nullptr, st));
  if (sh.masked0 && sh.n0 > 0) UMD_TRY(mask_argsort(a->mask_noise0, sh.n0, L, sh.keep0, ss.ids_shuffle, ss.ids_restore, nullptr, st));
  if (sh.masked1 && sh.n1 > 0)
    UMD_TRY(mask_argsort(a->mask_noise1, sh.n1, L, sh.keep1, ss.ids_shuffle + static_cast<long long>(sh.n0) * L,
                         ss.ids_restore + static_cast<long long>(sh.n0) * L, nullptr, st));
  umd_io io;
  memset(&io, 0, sizeof(io));
  io.image = ss.model_in; io.t = ss.t_model; io.labels = c.num_classes > 0 ? ss.labels_model : nullptr;
  io.ids_shuffle = ss.ids_shuffle; io.ids_restore = ss.ids_restore; io.x0 = a->image; io.noise = a->noise; io.loss = loss;
  UMD_TRY(engine_forward(&c, &sh, a->offsets, a->opt.params, a->opt.params_bf16, &io, a->workspace, ws_engine, 1, st));
  ReduceCtx red{a, static_cast<Comm*>(a->comm), st, UMD_OK};
  const bool reduce = a->comm != nullptr && comm_world(red.comm) > 1;
  UMD_TRY(engine_backward(&c, &sh, a->offsets, a->opt.params, a->opt.params_bf16, grads, &io, a->workspace, ws_engine,
                          reduce ? reduce_cb : nullptr, &red, st));
  UMD_TRY(red.rc);
  if (reduce) UMD_TRY(comm_join(red.comm, st));   // implicit GSPMD all-reduce of train_ae.py:364
  if (optimise) {
    umd_adamw_args o = a->opt;
    o.grads = grads;
    o.measurements = a->opt.measurements + 1;
    UMD_TRY(umd_adamw_step(&o, st));
    UMD_CHECK_CUDA(cudaMemcpyAsync(a->opt.measurements, loss, sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return UMD_OK;
}

extern "C" int umd_qsample(const float* x0, const float* noise, const int* t, const float* sa, const float* sb, int n,
                           int per_sample, float* out, umd_stream_t stream) {
  return qsample(x0, noise, t, sa, sb, n, per_sample, out, static_cast<cudaStream_t>(stream));
}
extern "C" int umd_ddim_step(const float* x, const float* pred, const float* noise, const int* t, const int* t_next,
                             const float* alphas_cumprod, const float* alphas_cumprod_prev,
                             const float* sqrt_recip_alphas_cumprod, const float* sqrt_recipm1_alphas_cumprod, int n, int hw,
                             int channels, int pred_channels, float eta, int use_cfg, float cfg_scale, int eps_pred, int clip_denoised,
                             float* sample, float* pred_xstart, umd_stream_t stream) {
  DdimArgs a;
  a.x = x; a.pred = pred; a.noise = noise; a.t = t; a.t_next = t_next;
  a.ac = alphas_cumprod; a.ac_prev = alphas_cumprod_prev;
  a.sqrt_recip_ac = sqrt_recip_alphas_cumprod; a.sqrt_recipm1_ac = sqrt_recipm1_alphas_cumprod;
  a.n = n; a.hw = hw; a.C = channels; a.pred_ld = pred_channels; a.eta = eta; a.cfg_scale = cfg_scale;
  a.use_cfg = use_cfg; a.eps_pred = eps_pred; a.clip_denoised = clip_denoised;
  a.sample = sample; a.pred_xstart = pred_xstart;
  return ddim_step(a, static_cast<cudaStream_t>(stream));
}
extern "C" int umd_mask_argsort(const float* noise, int n, int L, int len_keep, int* ids_shuffle, int* ids_restore,
                                float* mask, umd_stream_t stream) {
  return mask_argsort(noise, n, L, len_keep, ids_shuffle, ids_restore, mask, static_cast<cudaStream_t>(stream));
}
static AttnArgs mk_attn(const void* qkv, void* out, float* lse, int n0, int s0, int n1, int s1, int H, int Dh) {
  AttnArgs a;
  a.qkv = static_cast<const __nv_bfloat16*>(qkv); a.out = static_cast<__nv_bfloat16*>(out); a.lse = lse;
  a.rm = ragged_rowmap(n0, s0, n1, s1); a.nsamples = n0 + n1; a.H = H; a.Dh = Dh; a.scale = 1.0f / sqrtf(static_cast<float>(Dh));
  return a;
}
static AttnBwdArgs mk_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int n0,
                               int s0, int n1, int s1, int H, int Dh) {
  AttnBwdArgs a;
  a.delta = nullptr;
  a.qkv = static_cast<const __nv_bfloat16*>(qkv); a.out = static_cast<const __nv_bfloat16*>(out);
  a.dout = static_cast<const __nv_bfloat16*>(dout); a.lse = lse; a.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  a.rm = ragged_rowmap(n0, s0, n1, s1); a.nsamples = n0 + n1; a.H = H; a.Dh = Dh; a.scale = 1.0f / sqrtf(static_cast<float>(Dh));
  return a;
}
extern "C" int umd_attention_fwd(const void* qkv, void* out, float* lse, int n0, int s0, int n1, int s1, int H, int Dh,
                                 umd_stream_t stream) {
  return attention_fwd(mk_attn(qkv, out, lse, n0, s0, n1, s1, H, Dh), static_cast<cudaStream_t>(stream));
}
extern "C" int umd_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int n0,
                                 int s0, int n1, int s1, int H, int Dh, umd_stream_t stream) {
  return attention_bwd(mk_attn_bwd(qkv, out, dout, lse, dqkv, n0, s0, n1, s1, H, Dh), static_cast<cudaStream_t>(stream));
}
extern "C" int umd_attention_bwd_delta(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv,
                                       int n0, int s0, int n1, int s1, int H, int Dh, umd_stream_t stream) {
  UMD_REQUIRE(delta != nullptr, "umd_attention_bwd_delta: delta is required");
  AttnBwdArgs a = mk_attn_bwd(qkv, nullptr, dout, lse, dqkv, n0, s0, n1, s1, H, Dh);
  a.delta = delta;
  return attention_bwd(a, static_cast<cudaStream_t>(stream));
}
extern "C" int umd_attention_fwd_simt(const void* qkv, void* out, float* lse, int n0, int s0, int n1, int s1, int H, int Dh,
                                      umd_stream_t stream) {
  return attention_fwd_simt(mk_attn(qkv, out, lse, n0, s0, n1, s1, H, Dh), static_cast<cudaStream_t>(stream));
}
extern "C" int umd_attention_bwd_simt(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                      int n0, int s0, int n1, int s1, int H, int Dh, umd_stream_t stream) {
  return attention_bwd_simt(mk_attn_bwd(qkv, out, dout, lse, dqkv, n0, s0, n1, s1, H, Dh), static_cast<cudaStream_t>(stream));
}
extern "C" int umd_ln_modulate_fwd(const float* x, const float* gamma, const float* beta, const float* shift,
                                   const float* scale, long long ldmod, int n0, int s0, int n1, int s1, int D, void* out,
                                   int out_is_bf16, float* mean, float* rstd, umd_stream_t stream) {
  LnFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.gamma = gamma; a.beta = beta; a.shift = shift; a.scale = scale; a.ldmod = ldmod;
  a.rm = ragged_rowmap(n0, s0, n1, s1); a.out = out; a.mean = mean; a.rstd = rstd; a.rows_out = n0 * s0 + n1 * s1;
  return ln_mod_fwd(a, D, out_is_bf16 != 0, static_cast<cudaStream_t>(stream));
}
extern "C" int umd_ln_modulate_bwd(const void* dy, int dy_is_bf16, const float* x, const float* mean, const float* rstd,
                                   const float* gamma, const float* beta, const float* scale, long long ldmod, int n0,
                                   int s0, int n1, int s1, int D, float* dx, int accumulate, float* dshift, float* dscale,
                                   long long ldd, float* dgamma, float* dbeta, umd_stream_t stream) {
  LnBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.dy = dy; a.x = x; a.mean = mean; a.rstd = rstd; a.gamma = gamma; a.beta = beta; a.scale = scale; a.ldmod = ldmod;
  a.rm = ragged_rowmap(n0, s0, n1, s1); a.dx = dx; a.dx_in = dx; a.accumulate = accumulate; a.dshift = dshift; a.dscale = dscale;
  a.ldd = ldd; a.dgamma = dgamma; a.dbeta = dbeta;
  return ln_mod_bwd(a, D, n0 + n1, dy_is_bf16 != 0, static_cast<cudaStream_t>(stream));
}
extern "C" int umd_ln_modulate_bwd_gated(const void* dy, int dy_is_bf16, const float* x, const float* mean, const float* rstd,
                                         const float* gamma, const float* beta, const float* scale, long long ldmod, int n0,
                                         int s0, int n1, int s1, int D, float* dx, int accumulate, float* dshift,
                                         float* dscale, long long ldd, float* dgamma, float* dbeta, void* dz_bf16,
                                         const void* z_bf16, const float* gate, long long ldgate, float* dgate,
                                         long long lddgate, float* dbias, umd_stream_t stream) {
  LnBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.dy = dy; a.x = x; a.mean = mean; a.rstd = rstd; a.gamma = gamma; a.beta = beta; a.scale = scale; a.ldmod = ldmod;
  a.rm = ragged_rowmap(n0, s0, n1, s1); a.dx = dx; a.dx_in = dx; a.accumulate = accumulate; a.dshift = dshift; a.dscale = dscale;
  a.ldd = ldd; a.dgamma = dgamma; a.dbeta = dbeta;
  a.g_dz = static_cast<__nv_bfloat16*>(dz_bf16); a.g_z = static_cast<const __nv_bfloat16*>(z_bf16); a.g_gate = gate;
  a.g_ldgate = ldgate; a.g_dgate = dgate; a.g_lddgate = lddgate; a.g_dbias = dbias;
  UMD_REQUIRE(dz_bf16 != nullptr, "umd_ln_modulate_bwd_gated: dz is required (use umd_ln_modulate_bwd without a gate stage)");
  return ln_mod_bwd(a, D, n0 + n1, dy_is_bf16 != 0, static_cast<cudaStream_t>(stream));
}
extern "C" int umd_cast_f32_to_bf16(const float* x, long long n, void* out, umd_stream_t stream) {
  return cast_bf16(x, n, out, static_cast<cudaStream_t>(stream));
}
// measurement aid: 0 = keep the bias-gradient passes on the caller's stream (per-kernel timings then do not overlap)
extern "C" void umd_debug_side_stream(int on) { umd::g_side_enabled = on ? 1 : 0; }
