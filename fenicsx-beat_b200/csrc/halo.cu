// K5: ghost-dof exchange between ranks (one rank per GPU).
//
// Replaces the MPI neighbourhood exchanges hidden in the reference's PDE step:
//   b.ghostUpdate(ADD, REVERSE)   src/beat/base_model.py:203-206  -> not needed: the device RHS is an SpMV
//                                   over OWNED rows with ghost columns of v_ already valid (K1 runs on
//                                   ghosts too, as the reference's ODE stage does, odesolver.py:189-190)
//   state.x.scatter_forward()     src/beat/base_model.py:242       -> halo_refresh(x)
//   MatMult scatter inside KSP    (PETSc, inside :236)              -> per-iteration halo of the CG vectors
//   MPI_Allreduce inside KSP      (PETSc, inside :236)              -> cross-rank reduction of the CG scalars
//
// Communication backend: NCCL point-to-point (ncclSend/ncclRecv grouped per neighbour) over NVLink.
#include <nccl.h>

#include <cstring>

#include "mono_ctx.h"

#define MONO_NCCL(c, call)                                                                     \
  do {                                                                                         \
    ncclResult_t r__ = (call);                                                                 \
    if (r__ != ncclSuccess)                                                                    \
      return mono_fail((c), MONO_E_NCCL, std::string(#call) + ": " + ncclGetErrorString(r__)); \
  } while (0)

namespace {

__global__ void pack_kernel(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ vec,
                            double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = vec[idx[i]];
}

}  // namespace

int halo_destroy(mono_ctx* c) {
  if (c->comm) {
    ncclCommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
  }
  return MONO_OK;
}

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
  MONO_NCCL(c, ncclGroupStart());
  for (int k = 0; k < c->n_nbr; ++k) {
    const int64_t ns = c->send_ptr[k + 1] - c->send_ptr[k];
    const int64_t nr = c->recv_ptr[k + 1] - c->recv_ptr[k];
    if (ns > 0) MONO_NCCL(c, ncclSend(c->send_buf + c->send_ptr[k], ns, ncclDouble, c->nbr_ranks[k], comm, c->stream));
    if (nr > 0)
      MONO_NCCL(c, ncclRecv(vec + c->n_owned + c->recv_ptr[k], nr, ncclDouble, c->nbr_ranks[k], comm, c->stream));
  }
  MONO_NCCL(c, ncclGroupEnd());
  c->launches++;
  return MONO_OK;
}

extern "C" {

int mono_comm_unique_id(void* id_out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
  ncclUniqueId id;
  ncclResult_t r = ncclGetUniqueId(&id);
  if (r != ncclSuccess) return mono_fail(nullptr, MONO_E_NCCL, std::string("ncclGetUniqueId: ") + ncclGetErrorString(r));
  memcpy(id_out128, &id, sizeof(id));
  return MONO_OK;
}

int mono_comm_init(mono_ctx* c, int nranks, int rank, const void* id128) {
  MONO_CHECK(c, nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  MONO_CHECK(c, c->comm == nullptr, "communicator already initialised");
  c->nranks = nranks;
  c->rank = rank;
  if (nranks == 1) return MONO_OK;
  MONO_CUDA(c, cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  MONO_NCCL(c, ncclCommInitRank(&comm, nranks, id, rank));
  c->comm = (ncclComm*)comm;
  MONO_CUDA(c, cudaMalloc(&c->red_buf, sizeof(double) * 8));
  return MONO_OK;
}

int mono_set_halo(mono_ctx* c, int n_nbr, const int32_t* nbr_ranks, const int32_t* send_ptr, const int32_t* send_idx,
                  const int32_t* recv_ptr) {
  MONO_CHECK(c, c->has_pde, "set matrices before the halo pattern");
  MONO_CHECK(c, n_nbr >= 0, "negative neighbour count");
  c->n_nbr = n_nbr;
  c->nbr_ranks.assign(nbr_ranks, nbr_ranks + n_nbr);
  c->send_ptr.assign(send_ptr, send_ptr + n_nbr + 1);
  c->recv_ptr.assign(recv_ptr, recv_ptr + n_nbr + 1);
  MONO_CHECK(c, c->recv_ptr[n_nbr] == c->n_ghost, "recv_ptr must cover exactly the ghost block");
  c->n_send = c->send_ptr[n_nbr];
  for (int64_t k = 0; k < c->n_send; ++k) MONO_CHECK(c, send_idx[k] >= 0 && send_idx[k] < c->n_owned, "send index is not an owned dof");
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
  return MONO_OK;
}

}  // extern "C"
