// comm.cu — spf_comm: the NCCL communicator behind the row-sharded build and the list-sharded
// query.  The reference is single-process (rayon only); this is the exchange step
// north_star adds: "an NCCL-over-NVLink allreduce of the per-centroid partial sums and counts each
// k-means iteration" and "per-GPU top-k merged at the end".
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "comm.cuh"

namespace spf {

namespace {
NcclApi g_api;
bool g_api_ok = false;
std::once_flag g_once;
char g_api_err[256] = "";

template <typename F>
bool sym(void* h, const char* name, F* out) {
  *out = reinterpret_cast<F>(dlsym(h, name));
  if (!*out) snprintf(g_api_err, sizeof(g_api_err), "libnccl has no symbol %s", name);
  return *out != nullptr;
}

void load_api() {
  // the copy already mapped into this process (same SONAME) wins; otherwise the system library
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    snprintf(g_api_err, sizeof(g_api_err), "libnccl.so.2 could not be loaded: %s", dlerror());
    return;
  }
  g_api_ok = sym(h, "ncclGetUniqueId", &g_api.GetUniqueId) && sym(h, "ncclCommInitRank", &g_api.CommInitRank) &&
             sym(h, "ncclCommDestroy", &g_api.CommDestroy) && sym(h, "ncclAllGather", &g_api.AllGather) &&
             sym(h, "ncclAllReduce", &g_api.AllReduce) && sym(h, "ncclSend", &g_api.Send) &&
             sym(h, "ncclRecv", &g_api.Recv) && sym(h, "ncclGroupStart", &g_api.GroupStart) &&
             sym(h, "ncclGroupEnd", &g_api.GroupEnd) && sym(h, "ncclGetErrorString", &g_api.GetErrorString);
}
}  // namespace

const NcclApi* nccl_api() {
  std::call_once(g_once, load_api);
  if (!g_api_ok) {
    fail(SPF_E_CUDA, "%s", g_api_err);
    return nullptr;
  }
  return &g_api;
}

int comm_allgather(spf_ctx* c, const spf_comm* comm, const void* send, void* recv, size_t bytes) {
  if (!comm || comm->world == 1) {
    if (send != recv) SPF_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return SPF_OK;
  }
  const NcclApi* api = nccl_api();
  if (!api) return SPF_E_CUDA;
  SPF_NCCL(api, api->AllGather(send, recv, bytes, ncclUint8, comm->nccl, c->stream));
  return SPF_OK;
}

int comm_alltoall(spf_ctx* c, const spf_comm* comm, const void* send, void* recv, size_t chunk) {
  if (!comm || comm->world == 1) {
    if (send != recv) SPF_CUDA(cudaMemcpyAsync(recv, send, chunk, cudaMemcpyDeviceToDevice, c->stream));
    return SPF_OK;
  }
  const NcclApi* api = nccl_api();
  if (!api) return SPF_E_CUDA;
  SPF_NCCL(api, api->GroupStart());
  for (int p = 0; p < comm->world; ++p) {
    SPF_NCCL(api, api->Send(static_cast<const char*>(send) + (size_t)p * chunk, chunk, ncclUint8, p, comm->nccl, c->stream));
    SPF_NCCL(api, api->Recv(static_cast<char*>(recv) + (size_t)p * chunk, chunk, ncclUint8, p, comm->nccl, c->stream));
  }
  SPF_NCCL(api, api->GroupEnd());
  return SPF_OK;
}

int comm_allreduce_min_f32(spf_ctx* c, const spf_comm* comm, float* buf, size_t count) {
  if (!comm || comm->world == 1 || count == 0) return SPF_OK;
  const NcclApi* api = nccl_api();
  if (!api) return SPF_E_CUDA;
  SPF_NCCL(api, api->AllReduce(buf, buf, count, ncclFloat32, ncclMin, comm->nccl, c->stream));
  return SPF_OK;
}

int comm_group_start(const spf_comm* comm) {
  if (!comm || comm->world == 1) return SPF_OK;
  const NcclApi* api = nccl_api();
  if (!api) return SPF_E_CUDA;
  SPF_NCCL(api, api->GroupStart());
  return SPF_OK;
}

int comm_group_end(const spf_comm* comm) {
  if (!comm || comm->world == 1) return SPF_OK;
  const NcclApi* api = nccl_api();
  if (!api) return SPF_E_CUDA;
  SPF_NCCL(api, api->GroupEnd());
  return SPF_OK;
}

}  // namespace spf

using namespace spf;

extern "C" {

int spf_comm_unique_id(uint8_t* id128) {
  if (!id128) return fail(SPF_E_INVALID, "spf_comm_unique_id: NULL argument");
  const NcclApi* api = nccl_api();
  if (!api) return SPF_E_CUDA;
  static_assert(sizeof(ncclUniqueId) == SPF_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  SPF_NCCL(api, api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return SPF_OK;
}

int spf_comm_create(spf_ctx* ctx, int world, int rank, const uint8_t* id128, spf_comm** out) {
  if (!ctx || !out) return fail(SPF_E_INVALID, "spf_comm_create: NULL argument");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return fail(SPF_E_INVALID, "rank %d outside world %d", rank, world);
  if (world > 1 && !id128) return fail(SPF_E_INVALID, "spf_comm_create: the unique id is required for world > 1");
  spf_comm* cm = new (std::nothrow) spf_comm();
  if (!cm) return fail(SPF_E_OOM, "out of host memory");
  cm->ctx = ctx;
  cm->world = world;
  cm->rank = rank;
  if (world > 1) {
    const NcclApi* api = nccl_api();
    if (!api) { delete cm; return SPF_E_CUDA; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) { delete cm; return fail(SPF_E_CUDA, "cudaSetDevice failed: %s", cudaGetErrorString(e)); }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r = api->CommInitRank(&cm->nccl, world, id, rank);
    if (r != ncclSuccess) {
      delete cm;
      return fail(SPF_E_CUDA, "ncclCommInitRank failed: %s", api->GetErrorString(r));
    }
  }
  *out = cm;
  return SPF_OK;
}

void spf_comm_destroy(spf_comm* cm) {
  if (!cm) return;
  if (cm->nccl) {
    const NcclApi* api = nccl_api();
    cudaSetDevice(cm->ctx->device);
    cudaStreamSynchronize(cm->ctx->stream);
    if (api) api->CommDestroy(cm->nccl);
  }
  delete cm;
}

int spf_comm_world(const spf_comm* cm) { return cm ? cm->world : 1; }
int spf_comm_rank(const spf_comm* cm) { return cm ? cm->rank : 0; }

}  // extern "C"
