// tmq_comm.cpp -- NCCL plumbing for the sharded path: half-spinor face exchange (ncclSend/ncclRecv over
// NVLink 5 / NVSwitch on a dedicated stream) and the scalar all-reduces of the CG.  NCCL is bound at run time
// with dlopen so that libtmq.so loads on a box without NCCL / without a GPU (symbol-export tests) and picks up
// the libnccl.so.2 already mapped by torch when driven from Python.  Replaces QUDA's comm_* layer over
// MPI/QMP (reference CMakeLists.txt:266-310) and the reductions' MPI_Allreduce.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
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
  SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(AllReduce, "ncclAllReduce") SYM(GroupStart, "ncclGroupStart")
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

// in-place sum of n doubles of the device scalar block across ranks
int comm_allreduce(tmq_ctx *c, double *d_ptr, int n, cudaStream_t st) {
  if (!c->comm || c->comm->nranks == 1) return 0;
  TMQ_NCCL(g_nccl.AllReduce(d_ptr, d_ptr, (size_t)n, ncclDouble, ncclSum, c->comm->comm, st));
  return 0;
}

}  // namespace tmq
