// comm.cuh — the inter-GPU exchange of the row-sharded build and the list-sharded query
// (SURVEY.md 8(e)): one process per GPU, NCCL over NVLink / NVSwitch.  libnccl is resolved at run
// time (dlopen), so the library itself keeps no link-time dependency on it: a single-GPU caller
// never needs NCCL, and inside a process that already holds an NCCL (e.g. next to torch) the same
// copy is used.
#pragma once
#include <nccl.h>

#include "common.cuh"

struct spf_comm {
  spf_ctx* ctx = nullptr;
  ncclComm_t nccl = nullptr;
  int world = 1, rank = 0;
};

namespace spf {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
// NULL (with the error set) when libnccl cannot be loaded.
const NcclApi* nccl_api();

#define SPF_NCCL(api, expr)                                                                   \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess)                                                                    \
      return ::spf::fail(SPF_E_CUDA, "%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(_r), \
                         __FILE__, __LINE__);                                                 \
  } while (0)

// All-gather of `bytes` bytes per rank on the context stream: recv holds world x bytes in rank
// order.  world == 1 (comm NULL): a device copy (or nothing when send == recv).
int comm_allgather(spf_ctx* c, const spf_comm* comm, const void* send, void* recv, size_t bytes);
// Personalised exchange: rank r sends bytes [p * chunk, (p+1) * chunk) of `send` to rank p and
// receives rank p's chunk for r into recv[p * chunk ...).
int comm_alltoall(spf_ctx* c, const spf_comm* comm, const void* send, void* recv, size_t chunk);
// In-place element-wise minimum of `count` floats over the ranks (no-op for one rank).
int comm_allreduce_min_f32(spf_ctx* c, const spf_comm* comm, float* buf, size_t count);
// Several exchanges issued between start and end travel as ONE aggregated NCCL operation (one launch
// instead of one per call).  No-ops for one rank.
int comm_group_start(const spf_comm* comm);
int comm_group_end(const spf_comm* comm);

}  // namespace spf
