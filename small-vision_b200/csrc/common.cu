// Host-side plumbing shared by all translation units: error string, launch counter,
// TMA descriptor encode through a run-time-resolved driver entry point.
#include "common.cuh"

#include <stdarg.h>

#include <mutex>

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

}  // namespace umd

extern "C" const char* umd_last_error(void) { return umd::g_err; }
extern "C" int umd_version(void) { return 100; }
extern "C" long long umd_launch_count(void) { return umd::g_launch_count; }
