// tmq_comm.cpp -- NCCL plumbing for the sharded path: half-spinor face exchange (ncclSend/ncclRecv over
// NVLink 5 / NVSwitch on a dedicated stream) and the scalar all-reduces of the CG.  NCCL is bound at run time
// with dlopen so that libtmq.so loads on a box without NCCL / without a GPU (symbol-export tests) and picks up
// the libnccl.so.2 already mapped by torch when driven from Python.  Replaces QUDA's comm_* layer over
// MPI/QMP (reference CMakeLists.txt:266-310) and the reductions' MPI_Allreduce.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <vector>
#include "tmq_internal.h"

namespace tmq {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.handle) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !g_nccl.handle; i++) g_nccl.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.handle) { set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return 1; }
#define SYM(field, name)                                                    \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.handle, name);                   \
  if (!g_nccl.field) { set_error("NCCL symbol %s missing", name); return 1; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return 0;
}

struct Comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
};

#define TMQ_NCCL(call)                                                                              \
  do {                                                                                              \
    ncclResult_t r__ = (call);                                                                      \
    if (r__ != ncclSuccess) { set_error("%s:%d NCCL error: %s", __FILE__, __LINE__, g_nccl.GetErrorString(r__)); return 1; } \
  } while (0)

int comm_unique_id(char id128[128]) {
  if (nccl_load()) return 1;
  ncclUniqueId id;
  TMQ_NCCL(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(id128, &id, 128);
  return 0;
}

int comm_init(tmq_ctx *c, const char id128[128], int nranks, int rank) {
  if (nccl_load()) return 1;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  Comm *cm = new Comm();
  cm->nranks = nranks; cm->rank = rank;
  TMQ_CUDA(cudaSetDevice(c->device));
  TMQ_NCCL(g_nccl.CommInitRank(&cm->comm, nranks, id, rank));
  c->comm = cm;
  return 0;
}

void comm_destroy(tmq_ctx *c) {
  if (c->comm) {
    if (c->comm->comm) g_nccl.CommDestroy(c->comm->comm);
    delete c->comm;
    c->comm = nullptr;
  }
}

// rank of the process at coord + delta along dim (t fastest, then z: rank = ((cx*gy+cy)*gz+cz)*gt+ct)
static int rank_of(const tmq_ctx *c, int dim, int delta) {
  int co[4] = {c->coord[0], c->coord[1], c->coord[2], c->coord[3]};
  co[dim] = (co[dim] + delta + c->grid[dim]) % c->grid[dim];
  return ((co[0] * c->grid[1] + co[1]) * c->grid[2] + co[2]) * c->grid[3] + co[3];
}

// Exchange all partitioned faces of one Dslash application on `st` (one NCCL group):
//   send_bwd -> rank-1 ; send_fwd -> rank+1 ; recv ghost[.][1] <- rank+1 ; recv ghost[.][0] <- rank-1
int comm_exchange(tmq_ctx *c, int pi, int prec, cudaStream_t st) {
  bool remote = false;
  for (int d = 0; d < 4; d++) remote = remote || (c->g.part[d] && c->grid[d] > 1);
  if (remote && !c->comm) { set_error("lattice is partitioned across ranks but tmq_comm_init was not called"); return 1; }
  // a dimension partitioned on a grid of extent 1 (tmq_force_partition, the reference's --partition
  // flag, qkxtm/QKXTM_util.cpp:1717-1720) wraps onto this rank: the exchange is two device copies
  for (int d = 0; d < 4; d++) {
    if (!c->g.part[d] || c->grid[d] > 1) continue;
    const size_t nbytes = (size_t)3 * c->g.face[d] * vec_bytes(prec);
    TMQ_CUDA(cudaMemcpyAsync(c->halo_recv[pi][d][1], c->halo_send[pi][d][0], nbytes, cudaMemcpyDeviceToDevice, st));
    TMQ_CUDA(cudaMemcpyAsync(c->halo_recv[pi][d][0], c->halo_send[pi][d][1], nbytes, cudaMemcpyDeviceToDevice, st));
  }
  if (!remote) return 0;
  TMQ_NCCL(g_nccl.GroupStart());
  for (int d = 0; d < 4; d++) {
    if (!c->g.part[d] || c->grid[d] == 1) continue;
    const size_t nbytes = (size_t)3 * c->g.face[d] * vec_bytes(prec);
    const int rm = rank_of(c, d, -1), rp = rank_of(c, d, +1);
    TMQ_NCCL(g_nccl.Send(c->halo_send[pi][d][0], nbytes, ncclChar, rm, c->comm->comm, st));
    TMQ_NCCL(g_nccl.Send(c->halo_send[pi][d][1], nbytes, ncclChar, rp, c->comm->comm, st));
    TMQ_NCCL(g_nccl.Recv(c->halo_recv[pi][d][1], nbytes, ncclChar, rp, c->comm->comm, st));
    TMQ_NCCL(g_nccl.Recv(c->halo_recv[pi][d][0], nbytes, ncclChar, rm, c->comm->comm, st));
  }
  TMQ_NCCL(g_nccl.GroupEnd());
  return 0;
}

// generic neighbour exchange along one partitioned dimension (one-off set-up traffic such as the gauge halo of the clover
// term): send_bwd -> rank-1, send_fwd -> rank+1, recv_from_fwd <- rank+1, recv_from_bwd <- rank-1; wraps locally on a grid of 1
int comm_sendrecv_dim(tmq_ctx *c, int dim, const void *send_bwd, const void *send_fwd, void *recv_from_fwd, void *recv_from_bwd,
                      size_t nbytes, cudaStream_t st) {
  if (c->grid[dim] == 1) {
    TMQ_CUDA(cudaMemcpyAsync(recv_from_fwd, send_bwd, nbytes, cudaMemcpyDeviceToDevice, st));
    TMQ_CUDA(cudaMemcpyAsync(recv_from_bwd, send_fwd, nbytes, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  if (!c->comm) { set_error("lattice is partitioned across ranks but tmq_comm_init was not called"); return 1; }
  const int rm = rank_of(c, dim, -1), rp = rank_of(c, dim, +1);
  TMQ_NCCL(g_nccl.GroupStart());
  TMQ_NCCL(g_nccl.Send(send_bwd, nbytes, ncclChar, rm, c->comm->comm, st));
  TMQ_NCCL(g_nccl.Send(send_fwd, nbytes, ncclChar, rp, c->comm->comm, st));
  TMQ_NCCL(g_nccl.Recv(recv_from_fwd, nbytes, ncclChar, rp, c->comm->comm, st));
  TMQ_NCCL(g_nccl.Recv(recv_from_bwd, nbytes, ncclChar, rm, c->comm->comm, st));
  TMQ_NCCL(g_nccl.GroupEnd());
  return 0;
}

// Peer-memory halo path: expose this rank's ghost arena to its neighbours with CUDA IPC and map theirs.  The
// 64-byte handles travel through an ncclAllGather; every rank then opens the arenas of its (at most four)
// neighbours, and an all-reduce makes the decision unanimous: if any mapping failed anywhere, every rank stays
// on the ncclSend/ncclRecv path.
int comm_setup_p2p(tmq_ctx *c) {
  if (!c->multi || !c->arena || !c->comm) return 0;
  bool remote = false;
  for (int d = 2; d < 4; d++) remote = remote || (c->g.part[d] && c->grid[d] > 1);
  if (!remote) return 0;
  const int n = c->comm->nranks, me = c->comm->rank;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  int ok = cudaIpcGetMemHandle(&mine, c->arena) == cudaSuccess ? 1 : 0;
  if (!ok) cudaGetLastError();
  char *d_all = nullptr;
  TMQ_CUDA(cudaMalloc((void **)&d_all, (size_t)64 * n));
  TMQ_CUDA(cudaMemcpyAsync(d_all + (size_t)64 * me, &mine, 64, cudaMemcpyHostToDevice, c->stream));
  TMQ_NCCL(g_nccl.AllGather(d_all + (size_t)64 * me, d_all, 64, ncclChar, c->comm->comm, c->stream));
  std::vector<cudaIpcMemHandle_t> all(n);
  TMQ_CUDA(cudaMemcpyAsync(all.data(), d_all, (size_t)64 * n, cudaMemcpyDeviceToHost, c->stream));
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  TMQ_CUDA(cudaFree(d_all));
  // map every rank's arena: neighbours carry the ghost faces, all ranks carry the scalar all-reduce mailboxes
  if (n > TMQ_MAX_RANKS) ok = 0;
  for (int r = 0; r < n && ok; r++) {
    if (r == me) { c->rank_arena[r] = c->arena; continue; }
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
    c->rank_arena[r] = (char *)p;
    c->ipc_opened.push_back(p);
  }
  for (int d = 2; d < 4 && ok; d++) {
    if (!c->g.part[d] || c->grid[d] == 1) continue;
    for (int dir = 0; dir < 2; dir++) c->peer_arena[d][dir] = c->rank_arena[rank_of(c, d, dir ? +1 : -1)];
  }
  // unanimous decision
  const double bad = ok ? 0.0 : 1.0;
  TMQ_CUDA(cudaMemcpyAsync(c->scal + SC_T3, &bad, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TMQ_NCCL(g_nccl.AllReduce(c->scal + SC_T3, c->scal + SC_T3, 1, ncclDouble, ncclSum, c->comm->comm, c->stream));
  double total = 0;
  TMQ_CUDA(cudaMemcpyAsync(&total, c->scal + SC_T3, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  if (total != 0.0) {
    for (int d = 2; d < 4; d++)
      if (c->g.part[d] && c->grid[d] > 1) c->peer_arena[d][0] = c->peer_arena[d][1] = nullptr;
    c->p2p = false;
  } else {
    c->p2p = c->opt_p2p != 0;
  }
  return 0;
}

// rendezvous: NCCL all-reduce of a scratch scalar (NCCL waits for every rank, however late) + host wait
int comm_barrier(tmq_ctx *c) {
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  TMQ_CUDA(cudaStreamSynchronize(c->comm_stream));
  if (!c->comm || c->comm->nranks == 1) return 0;
  TMQ_NCCL(g_nccl.AllReduce(c->scal + SC_BARRIER, c->scal + SC_BARRIER, 1, ncclDouble, ncclSum, c->comm->comm, c->stream));
  TMQ_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// in-place sum of n doubles of the device scalar block across ranks
int comm_allreduce(tmq_ctx *c, double *d_ptr, int n, cudaStream_t st) {
  if (!c->comm || c->comm->nranks == 1) return 0;
  if (c->p2p && n <= 4 && d_ptr >= c->scal && d_ptr + n <= c->scal + SC_COUNT) {
    // peer-memory all-reduce: one 32-thread launch, NVLink stores into every rank's mailbox
    P2PRed R;
    memset(&R, 0, sizeof(R));
    R.scal = c->scal; R.slot = (int)(d_ptr - c->scal); R.n = n;
    R.rank = c->comm->rank; R.nranks = c->comm->nranks; R.seq = ++c->red_seq;
    R.err = c->scal + SC_ERR;
    R.timeout_ns = (unsigned long long)c->opt_halo_timeout_ms * 1000000ull;
    R.cg_iter = c->cg_iter_cur;
    R.cg_stop = (c->cg_iter_cur > 0 && (R.slot == SC_R2_0 || R.slot == SC_R2_1) && n == 1) ? 1 : 0;
    for (int r = 0; r < R.nranks; r++) {
      R.mbox[r] = (double *)(c->rank_arena[r] + c->arena_layout.mbox);
      R.mflag[r] = (unsigned int *)(c->rank_arena[r] + c->arena_layout.mflag);
    }
    TMQ_CUDA(p2p_allreduce(R, st));
    c->launches++;
    return 0;
  }
  TMQ_NCCL(g_nccl.AllReduce(d_ptr, d_ptr, (size_t)n, ncclDouble, ncclSum, c->comm->comm, st));
  return 0;
}

}  // namespace tmq
