// K5: everything that crosses ranks (one rank per GPU, one process per rank).
//
// Replaces the MPI traffic hidden in the reference's PDE step:
//   b.ghostUpdate(ADD, REVERSE)   src/beat/base_model.py:203-206  -> not needed: the device RHS is an SpMV
//                                   over OWNED rows with ghost columns of v_ already valid (K1 runs on
//                                   ghosts too, as the reference's ODE stage does, odesolver.py:189-190)
//   MatMult scatter inside KSP    (PETSc, inside :236)              -> per-iteration exchange of the CG vector
//   MPI_Allreduce inside KSP      (PETSc, inside :236)              -> cross-rank sums of the CG scalars
//   state.x.scatter_forward()     src/beat/base_model.py:242       -> ghost refresh of the solution
//
// Data path: NOT a collective library call.  Every rank exports ONE allocation (the two tagged exchange
// buffers, the x landing zone and the cross-rank reduction records, pde_kernels.cu) with CUDA IPC; the
// neighbours map it, and the persistent PDE kernel stores boundary values and partial sums straight into
// the peers' memory over NVLink ({value, generation} in one 16-byte store), so the halo exchange and the
// all-reduce happen inside the solve, element by element, overlapped with the SpMV of the interior rows.
// This file does the set-up for that: it exchanges the IPC handles and the ghost-block layout of all ranks
// (one ncclAllGather at mono_set_halo time) and builds the send table.  NCCL is only the bootstrap (and the
// reference exchange the tests compare the in-kernel path with, halo_refresh); it is loaded with dlopen at
// mono_comm_init so that the library has no link-time NCCL dependency (the process may already hold torch's
// bundled NCCL, which must be the one that is used).
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "mono_ctx.h"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  std::string error;
};

NcclApi g_nccl;

// Resolution order: a libnccl.so.2 the process already holds (torch's bundled one when the caller imported
// torch), then $MONO_NCCL_LIB, then the default search path.
bool nccl_load() {
  if (g_nccl.handle) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) {
    const char* env = getenv("MONO_NCCL_LIB");
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    g_nccl.error = std::string("cannot load libnccl.so.2: ") + dlerror();
    return false;
  }
  bool ok = true;
  auto sym = [&](const char* name) {
    void* p = dlsym(h, name);
    if (!p) {
      ok = false;
      g_nccl.error = std::string("libnccl.so.2 lacks ") + name;
    }
    return p;
  };
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
  g_nccl.CommAbort = reinterpret_cast<decltype(g_nccl.CommAbort)>(sym("ncclCommAbort"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
  g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(sym("ncclGroupStart"));
  g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(sym("ncclGroupEnd"));
  g_nccl.Send = reinterpret_cast<decltype(g_nccl.Send)>(sym("ncclSend"));
  g_nccl.Recv = reinterpret_cast<decltype(g_nccl.Recv)>(sym("ncclRecv"));
  g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
  if (!ok) return false;
  g_nccl.handle = h;
  return true;
}

#define MONO_NCCL(c, call)                                                                              \
  do {                                                                                                  \
    ncclResult_t r__ = (call);                                                                          \
    if (r__ != ncclSuccess)                                                                             \
      return mono_fail((c), MONO_E_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r__));      \
  } while (0)

__global__ void pack_kernel(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ vec,
                            double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = vec[idx[i]];
}

// what every rank tells every other rank at mono_set_halo time
struct PeerMeta {
  cudaIpcMemHandle_t handle;   // of the rank's exchange allocation
  int64_t n_owned, n_ghost;
  int64_t nb, nl, off_xg, off_xrecs, n_recs;  // layout of that allocation, in 16-byte records: nb buffers of nl, ...
  int64_t ghost_off[kMaxRanks];  // where the ghosts owned by rank q start in this rank's ghost block (-1: none)
  int64_t ghost_cnt[kMaxRanks];
};

}  // namespace

// A barrier over the ranks on the context's stream (a 4-byte all-reduce), bounded in time: a peer that died must not
// hang the teardown of the survivors.  Returns false when it timed out or could not be issued.
static bool comm_barrier(mono_ctx* c, double timeout_s) {
  if (!c->comm || c->nranks <= 1) return true;
  int* d = nullptr;
  if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) return false;
  cudaMemsetAsync(d, 0, sizeof(int), c->stream);
  bool ok = g_nccl.AllReduce(d, d, 1, ncclInt, ncclSum, (ncclComm_t)c->comm, c->stream) == ncclSuccess;
  if (ok) {
    const auto t0 = std::chrono::steady_clock::now();
    while (cudaStreamQuery(c->stream) == cudaErrorNotReady) {
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) {
        ok = false;
        break;
      }
      std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
    (void)cudaGetLastError();
  }
  if (ok) cudaFree(d);  // (a timed-out collective still owns the buffer: leak 4 bytes rather than free under it)
  return ok;
}

// Collective teardown of the peer mappings: every rank's exchange allocation is mapped by its peers, whose persistent
// kernels store into it.  Order: (1) my own work is done (caller synchronised the stream), (2) barrier: nobody launches
// another store into my buffers, (3) close my mappings of the peers, (4) barrier: nobody still maps my allocation,
// and only then does the caller free it.  With a dead peer the barriers time out and the teardown proceeds.
int halo_destroy(mono_ctx* c) {
  const bool collective = c->peers_ready && c->comm != nullptr && c->nranks > 1;
  double tmo = 10.0;
  if (const char* e = getenv("MONO_TEARDOWN_TIMEOUT_S")) tmo = std::max(0.0, atof(e));
  bool alive = collective && tmo > 0.0 && comm_barrier(c, tmo);
  for (int q = 0; q < kMaxRanks; ++q) {
    if (c->peer_base[q]) cudaIpcCloseMemHandle(c->peer_base[q]);
    c->peer_base[q] = nullptr;
    c->peer_xrecs[q] = nullptr;
  }
  c->peers_ready = false;
  if (alive) alive = comm_barrier(c, tmo);
  if (c->comm) {
    // a communicator with a collective that never completed (dead peer) must be aborted: destroying it would wait
    if (collective && !alive && tmo > 0.0)
      g_nccl.CommAbort((ncclComm_t)c->comm);
    else
      g_nccl.CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
  }
  return MONO_OK;
}

// Reference exchange through NCCL point-to-point (tests compare the in-kernel exchange against it; host-set
// vectors whose ghosts the caller did not fill can be completed with it).
int halo_refresh(mono_ctx* c, double* vec) {
  if (c->nranks <= 1 || c->n_nbr == 0) return MONO_OK;
  if (!c->comm) return mono_fail(c, MONO_E_INVALID, "multi-rank context without communicator (mono_comm_init)");
  if (c->n_send > 0) {
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>((c->n_send + threads - 1) / threads, (int64_t)c->n_sm * 4);
    pack_kernel<<<blocks, threads, 0, c->stream>>>(c->n_send, c->send_idx_dev, vec, c->send_buf);
    c->launches++;
    MONO_CUDA(c, cudaGetLastError());
  }
  ncclComm_t comm = (ncclComm_t)c->comm;
  MONO_NCCL(c, g_nccl.GroupStart());
  for (int k = 0; k < c->n_nbr; ++k) {
    const int64_t ns = c->send_ptr[k + 1] - c->send_ptr[k];
    const int64_t nr = c->recv_ptr[k + 1] - c->recv_ptr[k];
    if (ns > 0) MONO_NCCL(c, g_nccl.Send(c->send_buf + c->send_ptr[k], ns, ncclDouble, c->nbr_ranks[k], comm, c->stream));
    if (nr > 0)
      MONO_NCCL(c, g_nccl.Recv(vec + c->n_owned + c->recv_ptr[k], nr, ncclDouble, c->nbr_ranks[k], comm, c->stream));
  }
  MONO_NCCL(c, g_nccl.GroupEnd());
  c->launches++;
  return MONO_OK;
}

int halo_allreduce_max(mono_ctx* c, double* v) {
  if (c->nranks <= 1) return MONO_OK;
  if (!c->comm) return mono_fail(c, MONO_E_INVALID, "multi-rank context without communicator (mono_comm_init)");
  double* d = nullptr;
  MONO_CUDA(c, cudaMalloc(&d, sizeof(double)));
  MONO_CUDA(c, cudaMemcpyAsync(d, v, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  MONO_NCCL(c, g_nccl.AllReduce(d, d, 1, ncclDouble, ncclMax, (ncclComm_t)c->comm, c->stream));
  c->launches++;
  MONO_CUDA(c, cudaMemcpyAsync(v, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(d);
  return MONO_OK;
}

extern "C" {

int mono_comm_unique_id(void* id_out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
  if (!id_out128) return mono_fail(nullptr, MONO_E_INVALID, "id_out128 is NULL");
  if (!nccl_load()) return mono_fail(nullptr, MONO_E_NCCL, g_nccl.error);
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return mono_fail(nullptr, MONO_E_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
  memcpy(id_out128, &id, sizeof(id));
  return MONO_OK;
}

int mono_comm_init(mono_ctx* c, int nranks, int rank, const void* id128) {
  MONO_CHECK(c, nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  MONO_CHECK(c, nranks <= kMaxRanks, "more ranks than one node holds GPUs (kMaxRanks)");
  MONO_CHECK(c, c->comm == nullptr, "communicator already initialised");
  c->nranks = nranks;
  c->rank = rank;
  if (nranks == 1) return MONO_OK;
  MONO_CHECK(c, id128 != nullptr, "id128 is NULL");
  if (!nccl_load()) return mono_fail(c, MONO_E_NCCL, g_nccl.error);
  MONO_CUDA(c, cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  MONO_NCCL(c, g_nccl.CommInitRank(&comm, nranks, id, rank));
  c->comm = (ncclComm*)comm;
  c->spin_timeout_ms = std::max(c->spin_timeout_ms, 30000.0);  // ranks reach their first step at different times
  if (const char* e = getenv("MONO_SPIN_TIMEOUT_MS")) c->spin_timeout_ms = std::max(1.0, atof(e));
  return MONO_OK;
}

int mono_set_halo(mono_ctx* c, int n_nbr, const int32_t* nbr_ranks, const int32_t* send_ptr, const int32_t* send_idx,
                  const int32_t* recv_ptr) {
  MONO_CHECK(c, c->has_pde, "set matrices before the halo pattern");
  MONO_CHECK(c, n_nbr >= 0, "negative neighbour count");
  MONO_CHECK(c, send_ptr && recv_ptr && (n_nbr == 0 || nbr_ranks), "halo arrays are NULL");
  MONO_CHECK(c, send_ptr[0] == 0 && recv_ptr[0] == 0, "send_ptr / recv_ptr must start at 0");
  for (int k = 0; k < n_nbr; ++k)
    MONO_CHECK(c, send_ptr[k + 1] >= send_ptr[k] && recv_ptr[k + 1] >= recv_ptr[k], "send_ptr / recv_ptr must be non-decreasing");
  MONO_CHECK(c, send_ptr[n_nbr] == 0 || send_idx, "send_idx is NULL");
  c->n_nbr = n_nbr;
  c->nbr_ranks.assign(nbr_ranks, nbr_ranks + n_nbr);
  c->send_ptr.assign(send_ptr, send_ptr + n_nbr + 1);
  c->recv_ptr.assign(recv_ptr, recv_ptr + n_nbr + 1);
  MONO_CHECK(c, c->recv_ptr[n_nbr] == c->n_ghost, "recv_ptr must cover exactly the ghost block");
  c->n_send = c->send_ptr[n_nbr];
  for (int64_t k = 0; k < c->n_send; ++k) MONO_CHECK(c, send_idx[k] >= 0 && send_idx[k] < c->n_owned, "send index is not an owned dof");
  for (int k = 0; k < n_nbr; ++k)
    MONO_CHECK(c, nbr_ranks[k] >= 0 && nbr_ranks[k] < c->nranks && nbr_ranks[k] != c->rank, "bad neighbour rank");
  if (c->send_idx_dev) cudaFree(c->send_idx_dev);
  if (c->send_buf) cudaFree(c->send_buf);
  c->send_idx_dev = nullptr;
  c->send_buf = nullptr;
  if (c->n_send > 0) {
    MONO_CUDA(c, cudaMalloc(&c->send_idx_dev, sizeof(int32_t) * c->n_send));
    MONO_CUDA(c, cudaMalloc(&c->send_buf, sizeof(double) * c->n_send));
    MONO_CUDA(c, cudaMemcpyAsync(c->send_idx_dev, send_idx, sizeof(int32_t) * c->n_send, cudaMemcpyHostToDevice, c->stream));
    MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  if (c->nranks <= 1) return MONO_OK;
  MONO_CHECK(c, c->comm != nullptr, "mono_comm_init first");

  // ---- tell everybody where our exchange allocation is and how our ghost block is laid out ---------------------
  PeerMeta mine{};
  MONO_CUDA(c, cudaIpcGetMemHandle(&mine.handle, c->exch));
  mine.n_owned = c->n_owned;
  mine.n_ghost = c->n_ghost;
  mine.nb = c->exch_nb;
  mine.nl = c->exch_nl;
  mine.off_xg = c->exch_off_xg;
  mine.off_xrecs = c->exch_off_xrecs;
  mine.n_recs = c->exch_recs;
  for (int q = 0; q < kMaxRanks; ++q) {
    mine.ghost_off[q] = -1;
    mine.ghost_cnt[q] = 0;
  }
  for (int k = 0; k < n_nbr; ++k) {
    mine.ghost_off[nbr_ranks[k]] = recv_ptr[k];
    mine.ghost_cnt[nbr_ranks[k]] = recv_ptr[k + 1] - recv_ptr[k];
  }
  std::vector<PeerMeta> all((size_t)c->nranks);
  char *dsend = nullptr, *drecv = nullptr;
  MONO_CUDA(c, cudaMalloc(&dsend, sizeof(PeerMeta)));
  MONO_CUDA(c, cudaMalloc(&drecv, sizeof(PeerMeta) * c->nranks));
  MONO_CUDA(c, cudaMemcpyAsync(dsend, &mine, sizeof(PeerMeta), cudaMemcpyHostToDevice, c->stream));
  MONO_NCCL(c, g_nccl.AllGather(dsend, drecv, sizeof(PeerMeta), ncclChar, (ncclComm_t)c->comm, c->stream));
  c->launches++;
  MONO_CUDA(c, cudaMemcpyAsync(all.data(), drecv, sizeof(PeerMeta) * c->nranks, cudaMemcpyDeviceToHost, c->stream));
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(dsend);
  cudaFree(drecv);

  // ---- map every other rank's allocation (reduction records: all ranks; ghost slots: neighbours) ---------------
  for (int q = 0; q < c->nranks; ++q) {
    if (q == c->rank) continue;
    if (c->peer_base[q]) {
      cudaIpcCloseMemHandle(c->peer_base[q]);
      c->peer_base[q] = nullptr;
    }
    cudaError_t e = cudaIpcOpenMemHandle(&c->peer_base[q], all[q].handle, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      c->peer_base[q] = nullptr;
      return mono_fail(c, MONO_E_CUDA,
                       "cudaIpcOpenMemHandle(rank " + std::to_string(q) + "): " + cudaGetErrorString(e) +
                           " (the multi-GPU path needs peer access between the GPUs of the node)");
    }
    c->peer_xrecs[q] = static_cast<SyncRec*>(c->peer_base[q]) + all[q].off_xrecs;
  }
  c->peer_xrecs[c->rank] = c->xrecs;

  // ---- send table: the i-th value we send to neighbour q lands in the i-th ghost q holds from us -----------------
  std::vector<int32_t> rows;
  std::vector<std::vector<void*>> dt((size_t)c->exch_nb);
  std::vector<void*> dx;
  for (int k = 0; k < n_nbr; ++k) {
    const int q = nbr_ranks[k];
    const int64_t cnt = send_ptr[k + 1] - send_ptr[k];
    if (all[q].ghost_cnt[c->rank] != cnt)
      return mono_fail(c, MONO_E_INVALID,
                       "halo mismatch: this rank sends " + std::to_string(cnt) + " dofs to rank " + std::to_string(q) +
                           " but that rank holds " + std::to_string(all[q].ghost_cnt[c->rank]) + " ghosts owned by it");
    if (all[q].nb != c->exch_nb)
      return mono_fail(c, MONO_E_INVALID, "ranks disagree on the preconditioner (number of exchange buffers): configure the "
                                          "PDE stage identically on every rank before mono_set_halo");
    SyncRec* base = static_cast<SyncRec*>(c->peer_base[q]);
    for (int64_t i = 0; i < cnt; ++i) {
      const int64_t g = all[q].ghost_off[c->rank] + i;  // index in q's ghost block
      rows.push_back(send_idx[send_ptr[k] + i]);
      for (int b = 0; b < c->exch_nb; ++b) dt[(size_t)b].push_back(base + b * all[q].nl + all[q].n_owned + g);
      dx.push_back(base + all[q].off_xg + g);
    }
  }
  int rc = pde_build_send_table(c, rows, dt, dx);
  if (rc) return rc;
  c->peers_ready = true;
  return MONO_OK;
}

/* Ghost refresh of the PDE solution through NCCL send/recv (reference exchange; the step itself refreshes the
 * ghosts inside the persistent kernel).  Collective over the neighbours. */
int mono_halo_refresh_nccl(mono_ctx* c) {
  MONO_CHECK(c, c->has_pde, "no PDE matrices");
  return halo_refresh(c, c->x);
}

}  // extern "C"
