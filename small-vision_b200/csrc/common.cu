// Host-side plumbing shared by all translation units: error string, launch counter,
// TMA descriptor encode through a run-time-resolved driver entry point.
#include "common.cuh"

#include <stdarg.h>

#include <mutex>
#include <vector>

namespace umd {

static thread_local char g_err[1024] = "";
long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t batch,
                   uint64_t ld_elems, uint64_t batch_stride_elems, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return UMD_ERR_CUDA;
  }
  if (batch <= 1) {
    batch = 1;
    batch_stride_elems = outer * ld_elems;  // any legal value; never used for addressing
  }
  if ((batch_stride_elems * 2) % 16 != 0) {
    // a batch stride that breaks TMA alignment only matters when batch > 1
    if (batch > 1) {
      set_error("make_tmap_bf16: batch stride %llu elements is not 16-byte aligned",
                (unsigned long long)batch_stride_elems);
      return UMD_ERR_INVALID;
    }
    batch_stride_elems = (batch_stride_elems + 7) & ~7ull;
  }
  cuuint64_t dims[3] = {inner, outer, batch};
  cuuint64_t strides[2] = {ld_elems * 2, batch_stride_elems * 2};
  cuuint32_t box[3] = {64, box_outer, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu batch=%llu ld=%llu bs=%llu box=%u",
              (int)r, base, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)batch,
              (unsigned long long)ld_elems, (unsigned long long)batch_stride_elems, box_outer);
    return UMD_ERR_CUDA;
  }
  return UMD_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- profiling scopes ---------------------------------------------------------------------
bool g_prof_on = false;
namespace {
struct ProfRec { cudaEvent_t e0, e1; int cat; double work; };
std::vector<ProfRec> g_recs;
std::vector<cudaEvent_t> g_pool;
size_t g_pool_used = 0;
int g_depth = 0;
constexpr size_t kMaxEvents = 1 << 17;
cudaEvent_t take_event() {
  if (g_pool_used == g_pool.size()) {
    if (g_pool.size() >= kMaxEvents) return nullptr;
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    g_pool.push_back(e);
  }
  return g_pool[g_pool_used++];
}
}  // namespace
void prof_open(int cat, double work, cudaStream_t st) {
  if (g_depth++ > 0) return;  // only the outermost scope records
  ProfRec r;
  r.e0 = take_event();
  r.e1 = take_event();
  r.cat = cat;
  r.work = work;
  if (!r.e0 || !r.e1) { r.e0 = r.e1 = nullptr; }
  else cudaEventRecord(r.e0, st);
  g_recs.push_back(r);
}
void prof_close(cudaStream_t st) {
  if (--g_depth > 0) return;
  if (!g_recs.empty() && g_recs.back().e1) cudaEventRecord(g_recs.back().e1, st);
}

}  // namespace umd

extern "C" void umd_profile_enable(int on) {
  umd::g_prof_on = on != 0;
  umd::g_depth = 0;
}
extern "C" int umd_profile_num_categories(void) { return umd::PC_COUNT; }
extern "C" const char* umd_profile_category_name(int c) {
  static const char* names[umd::PC_COUNT] = {"gemm", "gemm_wgrad", "attention_fwd", "attention_bwd", "ln_modulate_fwd",
                                            "ln_modulate_bwd", "gate_bwd", "colsum", "optimizer"};
  return (c >= 0 && c < umd::PC_COUNT) ? names[c] : "?";
}
// Sums the recorded scopes per category (the caller has synchronised the device) and clears them.
extern "C" int umd_profile_read(float* ms, double* work, long long* scopes, int ncat) {
  using namespace umd;
  for (int i = 0; i < ncat; ++i) { ms[i] = 0.f; work[i] = 0.0; scopes[i] = 0; }
  int dropped = 0;
  for (const ProfRec& r : g_recs) {
    if (r.cat < 0 || r.cat >= ncat) continue;
    float t = 0.f;
    if (!r.e0 || cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) { ++dropped; continue; }
    ms[r.cat] += t;
    work[r.cat] += r.work;
    scopes[r.cat] += 1;
  }
  g_recs.clear();
  g_pool_used = 0;
  cudaGetLastError();
  return dropped;
}

extern "C" const char* umd_last_error(void) { return umd::g_err; }
extern "C" int umd_version(void) { return 100; }
extern "C" long long umd_launch_count(void) { return umd::g_launch_count; }
