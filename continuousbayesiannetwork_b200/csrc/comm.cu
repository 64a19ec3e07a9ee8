// The one collective of the path: the int64 sum of the per-GPU count tables over NVLink.
// The reference has no distributed code at all; this is what a sharded BruteForce._fit
// (cbn/parameter_learning/brute_force.py:42-43: counts, then counts / counts.sum()) needs between counting and
// normalising.  NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy the host process already loaded, e.g.
// PyTorch's, or the system library for a C-only caller), so the library itself has no link-time dependency on it and
// still loads on a machine without NCCL; the entry points then return CBN_ERR_UNSUPPORTED.
#include <dlfcn.h>
#include <string.h>

#include <new>

#include "common.cuh"

namespace {
// the few NCCL declarations this file needs (nccl.h 2.x ABI: enums and the 128-byte unique id are stable)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[CBN_COMM_ID_BYTES]; } ncclUniqueId;
constexpr int NCCL_INT64 = 4, NCCL_SUM = 0;

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {getenv("CBN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
      api.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(api.handle, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(api.handle, "ncclCommInitRank");
      api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.handle, "ncclCommDestroy");
      api.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
      api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
    }
  }
  return api;
}
}  // namespace

struct cbn_comm {
  int device = 0;
  int n_ranks = 1, rank = 0;
  ncclComm_t comm = nullptr;
};

extern "C" int cbn_comm_unique_id(uint8_t* id_out) {
  if (!id_out) return cbn_fail(nullptr, CBN_ERR_INVALID, "cbn_comm_unique_id: id_out is NULL");
  NcclApi& n = nccl();
  if (!n.ok) return cbn_fail(nullptr, CBN_ERR_UNSUPPORTED, "cbn_comm_unique_id: libnccl.so.2 could not be loaded (%s)", dlerror() ? "dlopen failed" : "symbols missing");
  ncclUniqueId id;
  const int rc = n.GetUniqueId(&id);
  if (rc != 0) return cbn_fail(nullptr, CBN_ERR_CUDA, "ncclGetUniqueId: %s", n.GetErrorString(rc));
  memcpy(id_out, id.internal, CBN_COMM_ID_BYTES);
  return CBN_OK;
}

extern "C" int cbn_comm_create(cbn_ctx* ctx, const uint8_t* id, int32_t n_ranks, int32_t rank, cbn_comm** out) {
  if (!ctx) return cbn_fail(nullptr, CBN_ERR_INVALID, "cbn_comm_create: ctx is NULL");
  if (!id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return cbn_fail(ctx, CBN_ERR_INVALID, "cbn_comm_create: bad argument");
  NcclApi& n = nccl();
  if (!n.ok) return cbn_fail(ctx, CBN_ERR_UNSUPPORTED, "cbn_comm_create: libnccl.so.2 could not be loaded");
  DeviceGuard g(ctx->device);
  // ncclCommInitRank binds the communicator to the CURRENT device
  int cur = -1;
  cudaGetDevice(&cur);
  if (cur != ctx->device) cudaSetDevice(ctx->device);
  cbn_comm* c = new (std::nothrow) cbn_comm();
  if (!c) return cbn_fail(ctx, CBN_ERR_NOMEM, "out of host memory");
  c->device = ctx->device; c->n_ranks = n_ranks; c->rank = rank;
  ncclUniqueId uid;
  memcpy(uid.internal, id, CBN_COMM_ID_BYTES);
  const int rc = n.CommInitRank(&c->comm, n_ranks, uid, rank);
  if (rc != 0) {
    delete c;
    return cbn_fail(ctx, CBN_ERR_CUDA, "ncclCommInitRank: %s", n.GetErrorString(rc));
  }
  *out = c;
  return CBN_OK;
}

extern "C" void cbn_comm_destroy(cbn_comm* comm) {
  if (!comm) return;
  NcclApi& n = nccl();
  if (n.ok && comm->comm) {
    DeviceGuard g(comm->device);
    n.CommDestroy(comm->comm);
  }
  delete comm;
}

extern "C" int cbn_comm_size(const cbn_comm* comm) { return comm ? comm->n_ranks : 0; }

extern "C" int cbn_counts_allreduce(cbn_ctx* ctx, cbn_comm* comm, long long* counts, int64_t n_cells, cbn_stream stream) {
  if (!ctx) return cbn_fail(nullptr, CBN_ERR_INVALID, "cbn_counts_allreduce: ctx is NULL");
  if (!comm || n_cells < 0 || (n_cells > 0 && !counts)) return cbn_fail(ctx, CBN_ERR_INVALID, "cbn_counts_allreduce: bad argument");
  if (n_cells == 0 || comm->n_ranks == 1) return CBN_OK;
  NcclApi& n = nccl();
  if (!n.ok) return cbn_fail(ctx, CBN_ERR_UNSUPPORTED, "cbn_counts_allreduce: libnccl.so.2 could not be loaded");
  DeviceGuard g(ctx->device);
  // integer addition is associative: whatever algorithm NCCL picks (ring, tree, NVLS in-switch reduction), the result is
  // bit-identical on every rank and for every number of ranks
  const int rc = n.AllReduce(counts, counts, (size_t)n_cells, NCCL_INT64, NCCL_SUM, comm->comm, (cudaStream_t)stream);
  if (rc != 0) return cbn_fail(ctx, CBN_ERR_CUDA, "ncclAllReduce: %s", n.GetErrorString(rc));
  return CBN_OK;
}
