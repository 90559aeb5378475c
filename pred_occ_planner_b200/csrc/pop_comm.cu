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
  // pipelined exchange (pop_exchange_trajs_pipelined): the allgather runs on its own stream between private
  // buffers, one call behind the cycle that consumes it
  cudaStream_t cs = nullptr;
  cudaEvent_t ev_send = nullptr, ev_snap = nullptr, ev_done[2] = {nullptr, nullptr};
  pop_traj_record* snap[2] = {nullptr, nullptr};    // this rank's records as sent, per parity of the call
  pop_traj_record* stage[2] = {nullptr, nullptr};   // all ranks' records as received
  bool pending[2] = {false, false};
  unsigned long long seq = 0;
};

void pop_comm_free(pop_handle* h) {
  CommDev* c = h->comm;
  if (!c) return;
  if (c->cs) cudaStreamSynchronize(c->cs);   // an allgather in flight finishes before its communicator goes
  if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
  for (int k = 0; k < 2; k++) {
    if (c->snap[k]) cudaFree(c->snap[k]);
    if (c->stage[k]) cudaFree(c->stage[k]);
    if (c->ev_done[k]) cudaEventDestroy(c->ev_done[k]);
  }
  if (c->ev_send) cudaEventDestroy(c->ev_send);
  if (c->ev_snap) cudaEventDestroy(c->ev_snap);
  if (c->cs) cudaStreamDestroy(c->cs);
  delete c;
  h->comm = nullptr;
}

static int comm_pipeline_alloc(pop_handle* h) {
  CommDev* c = h->comm;
  if (c->cs) return POP_OK;
  const pop_config& cfg = h->cfg;
  bool ok = cudaStreamCreateWithFlags(&c->cs, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->ev_send, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming) == cudaSuccess;
  for (int k = 0; k < 2 && ok; k++) {
    ok = ok && cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&c->snap[k], sizeof(pop_traj_record) * cfg.n_agents_local) == cudaSuccess;
    ok = ok && cudaMalloc(&c->stage[k], sizeof(pop_traj_record) * cfg.n_agents_total) == cudaSuccess;
  }
  if (!ok) {
    cudaGetLastError();
    return pop_fail(h, POP_E_CUDA, "pipelined exchange: stream / event / buffer allocation failed");
  }
  return POP_OK;
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

int pop_exchange_trajs_pipelined(pop_handle* h) {
  if (!h) return POP_E_INVALID;
  PopDeviceGuard guard(h);
  const pop_config& c = h->cfg;
  if (!h->comm || h->comm->world == 1 || !h->comm->comm) return pop_exchange_trajs(h);
  CommDev* cd = h->comm;
  int rc = comm_pipeline_alloc(h);
  if (rc) return rc;
  const int nl = c.n_agents_local, nt = c.n_agents_total, base = c.agent_id_base;
  const int cur = static_cast<int>(cd->seq & 1), prev = cur ^ 1;
  // 1. this rank's own records are visible to its next map update at once
  if ((rc = pop_traj_publish_local(h))) return rc;
  // 2. the other ranks' records of the previous call (its allgather has had a whole cycle to complete)
  if (cd->pending[prev]) {
    POP_CUDA(h, cudaStreamWaitEvent(h->stream, cd->ev_done[prev], 0));
    if (base > 0)
      POP_CUDA(h, cudaMemcpyAsync(h->d_traj_all, cd->stage[prev], sizeof(pop_traj_record) * base, cudaMemcpyDeviceToDevice, h->stream));
    if (base + nl < nt)
      POP_CUDA(h, cudaMemcpyAsync(h->d_traj_all + base + nl, cd->stage[prev] + base + nl,
                                  sizeof(pop_traj_record) * (nt - base - nl), cudaMemcpyDeviceToDevice, h->stream));
    cd->pending[prev] = false;
  }
  // 3. snapshot what this cycle published (the next cycle overwrites the send buffer long before every rank has
  //    arrived at this allgather) and start the allgather on the exchange stream
  POP_CUDA(h, cudaEventRecord(cd->ev_send, h->stream));
  POP_CUDA(h, cudaStreamWaitEvent(cd->cs, cd->ev_send, 0));
  POP_CUDA(h, cudaMemcpyAsync(cd->snap[cur], h->d_traj_send, sizeof(pop_traj_record) * nl, cudaMemcpyDeviceToDevice, cd->cs));
  POP_CUDA(h, cudaEventRecord(cd->ev_snap, cd->cs));
  POP_CUDA(h, cudaStreamWaitEvent(h->stream, cd->ev_snap, 0));
  const ncclResult_t r = nccl().AllGather(cd->snap[cur], cd->stage[cur], sizeof(pop_traj_record) * nl, ncclUint8, cd->comm, cd->cs);
  if (r != ncclSuccess) return pop_fail(h, POP_E_CUDA, "ncclAllGather failed: %s", nccl().GetErrorString(r));
  POP_CUDA(h, cudaEventRecord(cd->ev_done[cur], cd->cs));
  cd->pending[cur] = true;
  cd->seq++;
  h->have_traj = true;
  return POP_OK;
}

}  // extern "C"
