// pop_comm.cu — the per-cycle trajectory exchange behind the C ABI (SURVEY §8b "Search / coordinator", §8e).
//
// The reference broadcasts every freshly optimised trajectory on "/broadcast_traj"
// (plan_manager/src/fake_plan_manager.cpp:364-399) and every peer keeps the last one it heard
// (traj_coordinator/src/particles.cpp:131-191).  Here: ONE ncclAllGather of n_agents_local fixed-size records per
// cycle, on the handle's stream, from the handle's send buffer (written by the cycle's k_finalize) straight into
// its receive table (read by the next map update's peer projection, k_peer_positions / k_peer_scatter).  The
// communicator belongs to the handle.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2"): a host process that already carries an NCCL — PyTorch's
// wheel brings its own — keeps a single instance, and libpop.so itself does not depend on NCCL for single-GPU use.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "pop_internal.h"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi a;
  if (a.lib) return a;
  a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!a.lib) return a;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.lib, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.lib, "ncclCommInitRank"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.lib, "ncclAllGather"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.lib, "ncclGetErrorString"));
  a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(a.lib, "ncclGetVersion"));
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString;
  return a;
}

}  // namespace

struct CommDev {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

void pop_comm_free(pop_handle* h) {
  if (!h->comm) return;
  if (h->comm->comm && nccl().ok) nccl().CommDestroy(h->comm->comm);
  delete h->comm;
  h->comm = nullptr;
}

extern "C" {

int pop_comm_unique_id(void* id_out) {
  if (!id_out) return POP_E_INVALID;
  NcclApi& a = nccl();
  if (!a.ok) return POP_E_UNSUPPORTED;
  ncclUniqueId id;
  if (a.GetUniqueId(&id) != ncclSuccess) return POP_E_CUDA;
  static_assert(sizeof(ncclUniqueId) == POP_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  std::memcpy(id_out, &id, sizeof(id));
  return POP_OK;
}

int pop_comm_init(pop_handle* h, const void* nccl_unique_id, int rank, int world) {
  if (!h || world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_unique_id)) return POP_E_INVALID;
  const pop_config& c = h->cfg;
  if (c.n_agents_local * world != c.n_agents_total || c.agent_id_base != rank * c.n_agents_local)
    return pop_fail(h, POP_E_INVALID, "rank %d of %d does not match the handle's block of the swarm (ids %d..%d of %d): the "
                    "allgather needs equal blocks in rank order", rank, world, c.agent_id_base,
                    c.agent_id_base + c.n_agents_local - 1, c.n_agents_total);
  pop_comm_free(h);
  h->comm = new CommDev();
  h->comm->rank = rank;
  h->comm->world = world;
  if (world == 1) return POP_OK;
  NcclApi& a = nccl();
  if (!a.ok) { pop_comm_free(h); return pop_fail(h, POP_E_UNSUPPORTED, "libnccl.so.2 could not be loaded: %s", dlerror()); }
  PopDeviceGuard guard(h);
  ncclUniqueId id;
  std::memcpy(&id, nccl_unique_id, sizeof(id));
  const ncclResult_t r = a.CommInitRank(&h->comm->comm, world, id, rank);
  if (r != ncclSuccess) {
    h->comm->comm = nullptr;
    pop_comm_free(h);
    return pop_fail(h, POP_E_CUDA, "ncclCommInitRank failed: %s", a.GetErrorString(r));
  }
  return POP_OK;
}

int pop_comm_info(const pop_handle* h, int* rank, int* world, int* nccl_version) {
  if (!h) return POP_E_INVALID;
  if (rank) *rank = h->comm ? h->comm->rank : 0;
  if (world) *world = h->comm ? h->comm->world : 1;
  if (nccl_version) { *nccl_version = 0; if (nccl().ok && nccl().GetVersion) nccl().GetVersion(nccl_version); }
  return POP_OK;
}

int pop_exchange_trajs(pop_handle* h) {
  if (!h) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  const pop_config& c = h->cfg;
  if (!h->comm || h->comm->world == 1 || !h->comm->comm) {
    if (c.n_agents_local != c.n_agents_total && !(h->comm && h->comm->world == 1))
      return pop_fail(h, POP_E_STATE, "this handle holds %d of %d agents: call pop_comm_init before pop_exchange_trajs",
                      c.n_agents_local, c.n_agents_total);
    return pop_traj_publish_local(h);
  }
  const ncclResult_t r = nccl().AllGather(h->d_traj_send, h->d_traj_all, sizeof(pop_traj_record) * c.n_agents_local, ncclUint8,
                                          h->comm->comm, h->stream);
  if (r != ncclSuccess) return pop_fail(h, POP_E_CUDA, "ncclAllGather failed: %s", nccl().GetErrorString(r));
  h->have_traj = true;
  return POP_OK;
}

}  // extern "C"
