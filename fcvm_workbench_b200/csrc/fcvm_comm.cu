// Multi-GPU plumbing: one process per GPU, element-partitioned mesh, NCCL over NVLink.
// The shared-node sums (internal force, SpMV halo) and the CG dot products are all-reduces on
// the context's stream, so they order with the kernels without host synchronisation.
//
// NCCL is bound at run time (dlopen) so that the library loads on machines without it and so
// that a process that already carries an NCCL (e.g. through torch) uses that same copy.
#include <dlfcn.h>
#include <nccl.h>

#include "fcvm_common.cuh"

using namespace fcvm;

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) =
      nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return FCVM_OK;
  const char *cands[4] = {getenv("FCVM_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
  void *h = nullptr;
  for (int i = 0; i < 3 && !h; i++)
    if (cands[i] && cands[i][0]) h = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("NCCL not found (set FCVM_NCCL_LIB to libnccl.so.2): %s", dlerror());
    return FCVM_E_NCCL;
  }
  g_nccl.handle = h;
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    set_error("NCCL library lacks a required symbol");
    g_nccl = NcclApi();
    return FCVM_E_NCCL;
  }
  return FCVM_OK;
}

#define FCVM_NCCL(call)                                                                       \
  do {                                                                                        \
    ncclResult_t r__ = (call);                                                                \
    if (r__ != ncclSuccess) {                                                                 \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                                 \
                g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error");           \
      return FCVM_E_NCCL;                                                                     \
    }                                                                                         \
  } while (0)

}  // namespace

extern "C" int fcvm_comm_unique_id(void *id128) {
  FCVM_CHECK(id128, FCVM_E_ARG, "fcvm_comm_unique_id: null buffer");
  FCVM_TRY(load_nccl());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  FCVM_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return FCVM_OK;
}

extern "C" int fcvm_comm_init(fcvm_ctx *c, const void *id128, int rank, int world) {
  FCVM_CHECK(c && id128 && world >= 1 && rank >= 0 && rank < world, FCVM_E_ARG, "fcvm_comm_init: bad argument");
  c->rank = rank;
  c->world = world;
  if (world == 1) return FCVM_OK;
  FCVM_TRY(load_nccl());
  FCVM_CUDA(cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm;
  FCVM_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  c->nccl_comm = (void *)comm;
  if (!c->comm_stream) {
    FCVM_CUDA(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    FCVM_CUDA(cudaEventCreateWithFlags(&c->ev_boundary, cudaEventDisableTiming));
    FCVM_CUDA(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
  }
  return FCVM_OK;
}

extern "C" int fcvm_comm_destroy_(fcvm_ctx *c) {
  if (c && c->nccl_comm && g_nccl.CommDestroy) {
    g_nccl.CommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
  }
  return FCVM_OK;
}

namespace fcvm {
int comm_allreduce_on(fcvm_ctx *c, double *dev, int64_t n, cudaStream_t st) {
  FCVM_CHECK(c && dev && n > 0, FCVM_E_ARG, "allreduce: bad argument");
  if (c->world <= 1) return FCVM_OK;
  FCVM_CHECK(c->nccl_comm, FCVM_E_NCCL, "allreduce: communicator not initialised (fcvm_comm_init)");
  FCVM_NCCL(g_nccl.AllReduce(dev, dev, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)c->nccl_comm, st));
  return FCVM_OK;
}

int fcvm_comm_allreduce_oop(fcvm_ctx *c, const double *send, double *recv, int64_t n) {
  FCVM_CHECK(c && send && recv && n > 0, FCVM_E_ARG, "allreduce: bad argument");
  if (c->world <= 1) {
    if (send != recv)
      FCVM_CUDA(cudaMemcpyAsync(recv, send, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    return FCVM_OK;
  }
  FCVM_CHECK(c->nccl_comm, FCVM_E_NCCL, "allreduce: communicator not initialised (fcvm_comm_init)");
  FCVM_NCCL(g_nccl.AllReduce(send, recv, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
  return FCVM_OK;
}
}  // namespace fcvm

extern "C" int fcvm_comm_allreduce_max(fcvm_ctx *c, double *dev, int64_t n) {
  FCVM_CHECK(c && dev && n > 0, FCVM_E_ARG, "allreduce: bad argument");
  if (c->world <= 1) return FCVM_OK;
  FCVM_CHECK(c->nccl_comm, FCVM_E_NCCL, "allreduce: communicator not initialised (fcvm_comm_init)");
  FCVM_NCCL(g_nccl.AllReduce(dev, dev, (size_t)n, ncclFloat64, ncclMax, (ncclComm_t)c->nccl_comm, c->stream));
  return FCVM_OK;
}

extern "C" int fcvm_comm_allreduce_sum(fcvm_ctx *c, double *dev, int64_t n) {
  return fcvm::fcvm_comm_allreduce_oop(c, dev, dev, n);
}
