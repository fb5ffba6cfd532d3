// comm.cu -- the one collective of the path: the SUM all-reduce of the encoder + projector gradients once per
// optimizer step (the reference itself is single-GPU, its accumulation semantics are
// REF/trainer.py:372-384). The communicator belongs to the context (b2s_handle); buckets are all-reduced on the
// context's communication stream as soon as the compute stream has produced them (cudaStreamWaitEvent on the events
// b2s_hubert_backward records per transformer layer), so the exchange of layer l runs under the backward of layers
// l-1 ... 0 and the compute stream only joins at the very end.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 -- the copy PyTorch already mapped when there is one), so the
// library keeps loading on hosts without NCCL or without a GPU; only these entry points need it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "../../include/b2s.h"
#include "b2s_common.cuh"

namespace b2s {
namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

const NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy already in the process (PyTorch's), if any
    if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (h == nullptr) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
  });
  return api;
}

int need_nccl() {
  if (!nccl().ok) {
    set_last_error("NCCL (libnccl.so.2) could not be loaded: %s", dlerror() ? dlerror() : "symbols missing");
    return B2S_ERR_UNSUPPORTED;
  }
  return B2S_OK;
}

#define B2S_NCCL_CHECK(expr)                                                                  \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess) {                                                                  \
      set_last_error("NCCL error %d (%s): %s", static_cast<int>(_r), nccl().GetErrorString(_r), #expr); \
      return B2S_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

static_assert(sizeof(ncclUniqueId) == B2S_COMM_ID_BYTES, "b2s.h must carry NCCL's unique-id size");

}  // namespace

int comm_destroy(Context& c) {
  if (c.comm != nullptr && nccl().ok) nccl().CommDestroy(reinterpret_cast<ncclComm_t>(c.comm));
  c.comm = nullptr;
  if (c.comm_done != nullptr) cudaEventDestroy(c.comm_done);
  c.comm_done = nullptr;
  if (c.comm_stream != nullptr) cudaStreamDestroy(c.comm_stream);
  c.comm_stream = nullptr;
  c.comm_rank = 0;
  c.comm_world = 1;
  return B2S_OK;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

int b2s_comm_unique_id(uint8_t* id) {
  B2S_REQUIRE(id != nullptr, "b2s_comm_unique_id: null pointer");
  int rc = need_nccl();
  if (rc != B2S_OK) return rc;
  ncclUniqueId u;
  B2S_NCCL_CHECK(nccl().GetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return B2S_OK;
}

int b2s_comm_init(const uint8_t* id, int32_t rank, int32_t world) {
  B2S_REQUIRE(id != nullptr && world >= 1 && rank >= 0 && rank < world, "b2s_comm_init: bad arguments");
  int rc = need_nccl();
  if (rc != B2S_OK) return rc;
  Context& c = ctx();
  comm_destroy(c);
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t comm = nullptr;
  B2S_NCCL_CHECK(nccl().CommInitRank(&comm, world, u, rank));
  c.comm = comm;
  c.comm_rank = rank;
  c.comm_world = world;
  {
    // highest priority: when an SM frees a slot, a waiting NCCL CTA is placed before the next compute block, so the
    // exchange keeps pace with the backward it hides behind instead of queueing behind whole compute grids
    int prio_lo = 0, prio_hi = 0;
    B2S_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    B2S_CUDA_CHECK(cudaStreamCreateWithPriority(&c.comm_stream, cudaStreamNonBlocking, prio_hi));
  }
  B2S_CUDA_CHECK(cudaEventCreateWithFlags(&c.comm_done, cudaEventDisableTiming));
  return B2S_OK;
}

int b2s_comm_world(int32_t* rank, int32_t* world) {
  Context& c = ctx();
  if (rank) *rank = c.comm_rank;
  if (world) *world = c.comm != nullptr ? c.comm_world : 1;
  return B2S_OK;
}

int b2s_comm_destroy(void) { return comm_destroy(ctx()); }

int b2s_allreduce_grads(float* flat, const int64_t* starts, const int64_t* ends, void* const* ready_events,
                        int32_t n_buckets, void* span_begin_event, void* span_end_event) {
  Context& c = ctx();
  B2S_REQUIRE(c.comm != nullptr, "b2s_allreduce_grads: no communicator on this context (b2s_comm_init)");
  B2S_REQUIRE(flat && starts && ends && n_buckets >= 0, "b2s_allreduce_grads: bad arguments");
  bool first = true;
  for (int i = 0; i < n_buckets; ++i) {
    const int64_t n = ends[i] - starts[i];
    if (n <= 0) continue;
    if (ready_events != nullptr && ready_events[i] != nullptr)
      B2S_CUDA_CHECK(cudaStreamWaitEvent(c.comm_stream, reinterpret_cast<cudaEvent_t>(ready_events[i]), 0));
    if (first && span_begin_event != nullptr)
      B2S_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(span_begin_event), c.comm_stream));
    first = false;
    float* p = flat + starts[i];
    B2S_NCCL_CHECK(nccl().AllReduce(p, p, static_cast<size_t>(n), ncclFloat, ncclSum,
                                    reinterpret_cast<ncclComm_t>(c.comm), c.comm_stream));
  }
  if (!first && span_end_event != nullptr)
    B2S_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(span_end_event), c.comm_stream));
  return B2S_OK;
}

int b2s_allreduce_join(void* compute_stream) {
  Context& c = ctx();
  B2S_REQUIRE(c.comm != nullptr, "b2s_allreduce_join: no communicator on this context (b2s_comm_init)");
  B2S_CUDA_CHECK(cudaEventRecord(c.comm_done, c.comm_stream));
  B2S_CUDA_CHECK(cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(compute_stream), c.comm_done, 0));
  return B2S_OK;
}

}  // extern "C"
