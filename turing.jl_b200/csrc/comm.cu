// comm.cu -- NCCL plumbing for the two multi-GPU modes:
//   * rows sharded: sum of the [D+1][chains] likelihood/gradient block per leapfrog, enqueued
//     on the engine's stream right between the row-chunk reduction and the prior/scatter kernel;
//   * chains sharded: no data-path collective; diagnostics combine on the host side.
// One process per GPU; the 128-byte ncclUniqueId travels through the host's process group
// (torch.distributed in bench.py / tests, Distributed.jl or MPI from Julia).
#include <nccl.h>

#include <cstring>

#include "engine.cuh"

namespace tnuts {

int comm_allreduce_sum(Engine *e, double *buf, size_t n) {
  if (!e->comm || e->world <= 1) return TNUTS_OK;
  ncclResult_t r = ncclAllReduce(buf, buf, n, ncclDouble, ncclSum, (ncclComm_t)e->comm, e->stream);
  if (r != ncclSuccess)
    return set_error(e, TNUTS_E_NCCL, std::string("ncclAllReduce: ") + ncclGetErrorString(r));
  return TNUTS_OK;
}

// one grouped exchange: send slab q to rank q (when send_elems[q] > 0); when do_recv, receive
// rank q's slab at recv + recv_off[q]
int comm_exchange_slabs(Engine *e, const double *const *send, const size_t *send_elems, bool do_recv,
                        double *recv, const size_t *recv_off, const size_t *recv_elems) {
  if (!e->comm) return set_error(e, TNUTS_E_NCCL, "comm_exchange_slabs: no communicator");
  ncclComm_t c = (ncclComm_t)e->comm;
  ncclResult_t r = ncclGroupStart();
  for (int q = 0; q < e->world && r == ncclSuccess; ++q) {
    if (send_elems[q] > 0) r = ncclSend(send[q], send_elems[q], ncclDouble, q, c, e->stream);
    if (r == ncclSuccess && do_recv && recv_elems[q] > 0)
      r = ncclRecv(recv + recv_off[q], recv_elems[q], ncclDouble, q, c, e->stream);
  }
  ncclResult_t r2 = ncclGroupEnd();
  if (r != ncclSuccess || r2 != ncclSuccess)
    return set_error(e, TNUTS_E_NCCL, std::string("comm_exchange_slabs: ") + ncclGetErrorString(r != ncclSuccess ? r : r2));
  return TNUTS_OK;
}

void comm_destroy(Engine *e) {
  if (e->comm) {
    ncclCommDestroy((ncclComm_t)e->comm);
    e->comm = nullptr;
  }
}

}  // namespace tnuts

using namespace tnuts;

extern "C" {

int tnuts_comm_unique_id(void *unique_id_128) {
  if (!unique_id_128) return TNUTS_E_ARG;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return TNUTS_E_NCCL;
  std::memcpy(unique_id_128, &id, sizeof(id));
  return TNUTS_OK;
}

int tnuts_comm_init(tnuts_engine *h, const void *unique_id_128, int32_t world, int32_t rank) {
  Engine *e = (Engine *)h;
  if (!e || !unique_id_128 || world < 1 || rank < 0 || rank >= world)
    return set_error(e, TNUTS_E_ARG, "tnuts_comm_init: bad arguments");
  if (e->comm) return set_error(e, TNUTS_E_ARG, "tnuts_comm_init: communicator already initialised");
  TN_CUDA(e, cudaSetDevice(e->device));
  ncclUniqueId id;
  std::memcpy(&id, unique_id_128, sizeof(id));
  ncclComm_t c;
  ncclResult_t r = ncclCommInitRank(&c, world, id, rank);
  if (r != ncclSuccess)
    return set_error(e, TNUTS_E_NCCL, std::string("ncclCommInitRank: ") + ncclGetErrorString(r));
  e->comm = (ncclComm *)c;
  e->world = world;
  e->rank = rank;
  return TNUTS_OK;
}

/* sum a host vector over all ranks (used for cross-GPU diagnostics sufficient statistics) */
int tnuts_comm_allreduce_host(tnuts_engine *h, double *values, int64_t n) {
  Engine *e = (Engine *)h;
  if (!e || !values || n < 0) return set_error(e, TNUTS_E_ARG, "tnuts_comm_allreduce_host: bad call");
  if (!e->comm || e->world <= 1 || n == 0) return TNUTS_OK;
  TN_CUDA(e, cudaSetDevice(e->device));
  double *d = nullptr;
  TN_CUDA(e, cudaMalloc((void **)&d, 8 * (size_t)n));
  cudaError_t s = cudaMemcpyAsync(d, values, 8 * (size_t)n, cudaMemcpyHostToDevice, e->stream);
  int rc = s == cudaSuccess ? comm_allreduce_sum(e, d, (size_t)n) : TNUTS_E_CUDA;
  if (rc == TNUTS_OK) {
    s = cudaMemcpyAsync(values, d, 8 * (size_t)n, cudaMemcpyDeviceToHost, e->stream);
    if (s == cudaSuccess) s = cudaStreamSynchronize(e->stream);
    if (s != cudaSuccess) rc = set_error(e, TNUTS_E_CUDA, cudaGetErrorString(s));
  }
  cudaFree(d);
  return rc;
}

}  // extern "C"
