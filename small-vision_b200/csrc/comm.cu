// Data-parallel exchange below the C ABI: a thin communicator over NCCL (resolved at run time, so that the library still
// loads on a machine without NCCL or without a GPU) with its own communication stream.  The reference declares its
// shardings and lets GSPMD insert the gradient all-reduce (big_vision/trainers/train_ae.py:159-170,287-290,364); here the
// step engine hands every finished gradient bucket to umd_comm (engine.cu, umd_train_step) while the rest of the backward
// pass runs.  A host that already owns an ncclComm_t (an XLA / JAX runtime) wraps it with umd_comm_from_nccl.
#include "common.cuh"
#include "kernels.cuh"

#include <dlfcn.h>
#include <nccl.h>   // types and enums only: every entry point is looked up with dlsym

namespace umd {

namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*CommCount)(const ncclComm_t, int*);
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*);
  const char* (*GetErrorString)(ncclResult_t);
  ncclResult_t (*GetVersion)(int*);
  bool ok;
};

const NcclApi* nccl_api() {
  static NcclApi api;
  static int state = 0;   // 0 untried, 1 ok, -1 failed
  if (state == 0) {
    // the copy torch has already mapped (its bundled wheel) wins, then whatever the loader finds
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    state = -1;
    if (h) {
      api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
      api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
      api.CommCount = reinterpret_cast<decltype(api.CommCount)>(dlsym(h, "ncclCommCount"));
      api.CommUserRank = reinterpret_cast<decltype(api.CommUserRank)>(dlsym(h, "ncclCommUserRank"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
      api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(h, "ncclGetVersion"));
      if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.CommCount && api.CommUserRank &&
          api.GetErrorString)
        state = 1;
    }
  }
  return state == 1 ? &api : nullptr;
}

#define UMD_CHECK_NCCL(api, expr)                                                                          \
  do {                                                                                                     \
    ncclResult_t _r = (expr);                                                                              \
    if (_r != ncclSuccess) {                                                                               \
      ::umd::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(_r));          \
      return UMD_ERR_CUDA;                                                                                 \
    }                                                                                                      \
  } while (0)

}  // namespace

struct Comm {
  ncclComm_t nccl;
  bool owned;
  int world, rank;
  cudaStream_t stream;     // all collectives of this communicator run here
  cudaEvent_t ev[32];      // fork events (compute -> comm), round robin
  int next;
  cudaEvent_t done;        // last collective (comm -> compute)
};

int comm_world(const Comm* c) { return c ? c->world : 1; }

// Mean all-reduce of buf[0, n) on the communicator's stream, ordered behind everything enqueued so far on `after`
// (and on `after2` when given: the engine's side stream).
int comm_allreduce_mean_after(Comm* c, float* buf, long long n, cudaStream_t after, cudaEvent_t after2) {
  const NcclApi* api = nccl_api();
  UMD_REQUIRE(api && c, "umd_comm: NCCL is not available");
  if (n <= 0) return UMD_OK;
  cudaEvent_t e = c->ev[c->next++ & 31];
  UMD_CHECK_CUDA(cudaEventRecord(e, after));
  UMD_CHECK_CUDA(cudaStreamWaitEvent(c->stream, e, 0));
  if (after2) UMD_CHECK_CUDA(cudaStreamWaitEvent(c->stream, after2, 0));
  UMD_CHECK_NCCL(api, api->AllReduce(buf, buf, static_cast<size_t>(n), ncclFloat, ncclAvg, c->nccl, c->stream));
  return UMD_OK;
}
// `stream` waits for every collective issued so far
int comm_join(Comm* c, cudaStream_t stream) {
  UMD_REQUIRE(c, "umd_comm: null communicator");
  UMD_CHECK_CUDA(cudaEventRecord(c->done, c->stream));
  UMD_CHECK_CUDA(cudaStreamWaitEvent(stream, c->done, 0));
  return UMD_OK;
}

static int comm_finish_init(Comm* c) {
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  UMD_CHECK_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));   // collectives first: they are short and gate the optimiser
  for (int i = 0; i < 32; ++i) UMD_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming));
  UMD_CHECK_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
  c->next = 0;
  return UMD_OK;
}

}  // namespace umd

using namespace umd;

extern "C" int umd_comm_available(void) { return nccl_api() ? 1 : 0; }

extern "C" int umd_comm_nccl_version(void) {
  const NcclApi* api = nccl_api();
  int v = 0;
  if (api && api->GetVersion) api->GetVersion(&v);
  return v;
}

extern "C" int umd_comm_unique_id(unsigned char* id128) {
  const NcclApi* api = nccl_api();
  UMD_REQUIRE(api, "umd_comm_unique_id: libnccl.so.2 could not be loaded");
  UMD_REQUIRE(id128, "umd_comm_unique_id: null buffer");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes in every NCCL 2 release");
  ncclUniqueId id;
  UMD_CHECK_NCCL(api, api->GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return UMD_OK;
}

extern "C" int umd_comm_init(int rank, int world, const unsigned char* id128, void** out) {
  const NcclApi* api = nccl_api();
  UMD_REQUIRE(api, "umd_comm_init: libnccl.so.2 could not be loaded");
  UMD_REQUIRE(out && id128 && world >= 1 && rank >= 0 && rank < world, "umd_comm_init: bad arguments (rank %d of %d)", rank, world);
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  Comm* c = new Comm();
  c->owned = true; c->world = world; c->rank = rank;
  ncclResult_t r = api->CommInitRank(&c->nccl, world, id, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank(rank %d of %d) -> %s", rank, world, api->GetErrorString(r));
    delete c;
    return UMD_ERR_CUDA;
  }
  int rc = comm_finish_init(c);
  if (rc != UMD_OK) { api->CommDestroy(c->nccl); delete c; return rc; }
  *out = c;
  return UMD_OK;
}

extern "C" int umd_comm_from_nccl(void* nccl_comm, void** out) {
  const NcclApi* api = nccl_api();
  UMD_REQUIRE(api, "umd_comm_from_nccl: libnccl.so.2 could not be loaded");
  UMD_REQUIRE(nccl_comm && out, "umd_comm_from_nccl: null argument");
  Comm* c = new Comm();
  c->owned = false; c->nccl = static_cast<ncclComm_t>(nccl_comm);
  if (api->CommCount(c->nccl, &c->world) != ncclSuccess || api->CommUserRank(c->nccl, &c->rank) != ncclSuccess) {
    set_error("umd_comm_from_nccl: not a valid ncclComm_t");
    delete c;
    return UMD_ERR_INVALID;
  }
  int rc = comm_finish_init(c);
  if (rc != UMD_OK) { delete c; return rc; }
  *out = c;
  return UMD_OK;
}

extern "C" int umd_comm_world(void* comm) { return comm ? static_cast<Comm*>(comm)->world : 1; }
extern "C" int umd_comm_rank(void* comm) { return comm ? static_cast<Comm*>(comm)->rank : 0; }

extern "C" int umd_comm_allreduce_mean(void* comm, float* buf, long long n, umd_stream_t stream) {
  UMD_REQUIRE(comm, "umd_comm_allreduce_mean: null communicator");
  Comm* c = static_cast<Comm*>(comm);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UMD_TRY(comm_allreduce_mean_after(c, buf, n, st, nullptr));
  return comm_join(c, st);
}

extern "C" int umd_comm_destroy(void* comm) {
  if (!comm) return UMD_OK;
  Comm* c = static_cast<Comm*>(comm);
  const NcclApi* api = nccl_api();
  cudaStreamSynchronize(c->stream);
  if (c->owned && api) api->CommDestroy(c->nccl);
  for (int i = 0; i < 32; ++i) cudaEventDestroy(c->ev[i]);
  cudaEventDestroy(c->done);
  cudaStreamDestroy(c->stream);
  delete c;
  return UMD_OK;
}
